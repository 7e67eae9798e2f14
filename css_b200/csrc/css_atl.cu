// Attention_Threshold_Loss -- the logit-space unsupervised loss that consumes the pseudo labels and
// confidences produced by the path (sole user of the `un` threshold; in cross_label it is supervised by the
// representation-space label / confidence).  Reference: generalframeworks/loss/loss.py:48-64.
//   weighting_b = #(conf_b >= thr) / #(label_b >= 0)
//   loss_p      = CE(pred[:, p], label_p), 0 on ignored pixels (label -1)
//   out         = mean over {p : loss_p > 0} of weighting_b * loss_p
// HBM-bound: forward reads pred once (C*4 B / crop pixel) and keeps the per-pixel logsumexp (4 B), backward reads pred
// once more and writes grad_pred once.  The reference makes ~6 passes over [B,C,H,W] tensors plus a masked_select.
#include "css_common.cuh"

#define ATL_THREADS 256

// counts layout per image: [0] #(loss > 0), [1] #(conf >= thr), [2] #(label >= 0)
// CT > 0: compile-time class count (the C logits of a pixel stay in registers: pred is read exactly once); CT == 0: any C.
template <int CT>
__global__ void __launch_bounds__(ATL_THREADS) atl_forward_kernel(const float* __restrict__ pred, const int64_t* __restrict__ label,
                                                                  const float* __restrict__ conf, float thr, int C, int HW,
                                                                  float* __restrict__ lse_out, float* __restrict__ partials,
                                                                  int32_t* __restrict__ counts) {
    css_pdl_enter();
    __shared__ float wsum[ATL_THREADS / 32];
    __shared__ int wcnt[3][ATL_THREADS / 32];
    const int b = blockIdx.y, p = blockIdx.x * ATL_THREADS + threadIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float loss = 0.f;
    int pos = 0, hi = 0, valid = 0;
    if (p < HW) {
        const float* x = pred + (size_t)b * C * HW + p;
        const long long lab = label[(size_t)b * HW + p];
        float lse, picked = 0.f;
        if (CT > 0) {
            float v[CT > 0 ? CT : 1];
#pragma unroll
            for (int c = 0; c < CT; ++c) v[c] = ldg_stream(x + (size_t)c * HW);
            float m = v[0];
#pragma unroll
            for (int c = 1; c < CT; ++c) m = fmaxf(m, v[c]);
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                s += expf(v[c] - m);
                if (c == lab) picked = v[c];
            }
            lse = m + logf(s);
        } else {
            float m = -INFINITY;
            for (int c = 0; c < C; ++c) m = fmaxf(m, __ldg(x + (size_t)c * HW));
            float s = 0.f;
            for (int c = 0; c < C; ++c) s += expf(__ldg(x + (size_t)c * HW) - m);
            lse = m + logf(s);
            if (lab >= 0 && lab < C) picked = __ldg(x + (size_t)lab * HW);
        }
        lse_out[(size_t)b * HW + p] = lse;
        valid = lab >= 0;
        if (valid && lab < C) loss = lse - picked;
        pos = loss > 0.f;
        if (!pos) loss = 0.f;
        hi = conf[(size_t)b * HW + p] >= thr;
    }
    // fixed-order block reduction (deterministic): shuffle tree, then the 8 warp sums in warp order
    float v = warp_sum(loss);
    const int npos = __popc(__ballot_sync(0xffffffffu, pos)), nhi = __popc(__ballot_sync(0xffffffffu, hi)),
              nvalid = __popc(__ballot_sync(0xffffffffu, valid));
    if (lane == 0) {
        wsum[warp] = v;
        wcnt[0][warp] = npos;
        wcnt[1][warp] = nhi;
        wcnt[2][warp] = nvalid;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        int c0 = 0, c1 = 0, c2 = 0;
#pragma unroll
        for (int i = 0; i < ATL_THREADS / 32; ++i) {
            t += wsum[i];
            c0 += wcnt[0][i];
            c1 += wcnt[1][i];
            c2 += wcnt[2][i];
        }
        partials[(size_t)b * gridDim.x + blockIdx.x] = t;
        atomicAdd(counts + 3 * b + 0, c0);
        atomicAdd(counts + 3 * b + 1, c1);
        atomicAdd(counts + 3 * b + 2, c2);
    }
}

// one CTA: per-image loss sums (fixed order), the scalar loss and the per-image backward scale  w_b / #(loss > 0)
__global__ void __launch_bounds__(256) atl_finalize_kernel(const float* __restrict__ partials, const int32_t* __restrict__ counts, int B,
                                                           int nblk, float* __restrict__ scale, float* __restrict__ loss) {
    css_pdl_enter();
    __shared__ float img_sum[256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int b = warp; b < B; b += 8) {
        float s = 0.f;
        for (int i = lane; i < nblk; i += 32) s += partials[(size_t)b * nblk + i];
        s = warp_sum(s);
        if (lane == 0) img_sum[b] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long total_pos = 0;
        for (int b = 0; b < B; ++b) total_pos += counts[3 * b];
        float acc = 0.f;
        for (int b = 0; b < B; ++b) {
            const float wgt = (float)counts[3 * b + 1] / (float)counts[3 * b + 2];     // may be nan/inf: only used where loss > 0
            scale[b] = wgt / (float)total_pos;
            if (counts[3 * b] > 0) acc += wgt * img_sum[b];
        }
        *loss = acc / (float)total_pos;                      // mean of an empty selection is nan, as in torch
    }
}

template <int CT>
__global__ void __launch_bounds__(ATL_THREADS) atl_backward_kernel(const float* __restrict__ grad_out, const float* __restrict__ pred,
                                                                   const int64_t* __restrict__ label, const float* __restrict__ lse,
                                                                   const float* __restrict__ scale, int C, int HW,
                                                                   float* __restrict__ grad_pred) {
    css_pdl_enter();
    const int b = blockIdx.y, p = blockIdx.x * ATL_THREADS + threadIdx.x;
    if (p >= HW) return;
    const float* x = pred + (size_t)b * C * HW + p;
    float* g = grad_pred + (size_t)b * C * HW + p;
    const long long lab = label[(size_t)b * HW + p];
    const float l = lse[(size_t)b * HW + p];
    const bool labelled = lab >= 0 && lab < C;
    const float k = __ldg(grad_out) * scale[b];
    if (CT > 0) {
        float v[CT > 0 ? CT : 1];
        float picked = 0.f;
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            v[c] = labelled ? ldg_stream(x + (size_t)c * HW) : 0.f;
            if (c == lab) picked = v[c];
        }
        const bool on = labelled && (l - picked) > 0.f;       // masked_select(loss > 0): no gradient elsewhere
#pragma unroll
        for (int c = 0; c < CT; ++c) g[(size_t)c * HW] = on ? k * (expf(v[c] - l) - (c == lab ? 1.f : 0.f)) : 0.f;
        return;
    }
    bool on = labelled;
    if (on) on = (l - __ldg(x + (size_t)lab * HW)) > 0.f;
    if (!on) {
        for (int c = 0; c < C; ++c) g[(size_t)c * HW] = 0.f;
        return;
    }
    for (int c = 0; c < C; ++c) {
        const float sm = expf(ldg_stream(x + (size_t)c * HW) - l);
        g[(size_t)c * HW] = k * (sm - (c == lab ? 1.f : 0.f));
    }
}

extern "C" int css_atl_blocks(int H, int W) { return (H * W + ATL_THREADS - 1) / ATL_THREADS; }

extern "C" int css_atl_forward(const float* pred, const int64_t* label, const float* conf, float threshold, int B, int C, int H,
                               int W, float* lse, float* partials, int32_t* counts, float* scale, float* loss, void* stream) {
    CSS_CHECK_ARG(pred && label && conf && lse && partials && counts && scale && loss, CSS_E_ARG, "css_atl_forward: null pointer");
    CSS_CHECK_ARG(B > 0 && B <= 256 && C > 0 && H > 0 && W > 0, CSS_E_ARG, "css_atl_forward: bad size (B must be in [1,256])");
    CSS_CHECK_ARG((long long)B * C * H * W < (1ll << 40) && (long long)H * W < (1ll << 31), CSS_E_SIZE, "css_atl_forward: too large");
    cudaStream_t st = (cudaStream_t)stream;
    const int HW = H * W, nblk = css_atl_blocks(H, W);
    cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(int32_t) * 3 * B, st);
    if (e != cudaSuccess) { css_set_error("css_atl_forward: memset: %s", cudaGetErrorString(e)); return (int)e; }
    if (C == 21) css_launch(atl_forward_kernel<21>, dim3(dim3(nblk, B)), dim3(ATL_THREADS), (size_t)(0), (cudaStream_t)(st), pred, label, conf, threshold, C, HW, lse, partials, counts);
    else if (C == 19) css_launch(atl_forward_kernel<19>, dim3(dim3(nblk, B)), dim3(ATL_THREADS), (size_t)(0), (cudaStream_t)(st), pred, label, conf, threshold, C, HW, lse, partials, counts);
    else css_launch(atl_forward_kernel<0>, dim3(dim3(nblk, B)), dim3(ATL_THREADS), (size_t)(0), (cudaStream_t)(st), pred, label, conf, threshold, C, HW, lse, partials, counts);
    css_launch(atl_finalize_kernel, dim3(1), dim3(256), (size_t)(0), (cudaStream_t)(st), partials, counts, B, nblk, scale, loss);
    CSS_CHECK_LAUNCH("css_atl_forward", 2);
    return 0;
}

extern "C" int css_atl_backward(const float* grad_out, const float* pred, const int64_t* label, const float* lse, const float* scale,
                                int B, int C, int H, int W, float* grad_pred, void* stream) {
    CSS_CHECK_ARG(grad_out && pred && label && lse && scale && grad_pred, CSS_E_ARG, "css_atl_backward: null pointer");
    CSS_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0, CSS_E_ARG, "css_atl_backward: bad size");
    const int HW = H * W;
    const dim3 grid(css_atl_blocks(H, W), B);
    cudaStream_t st = (cudaStream_t)stream;
    if (C == 21) css_launch(atl_backward_kernel<21>, dim3(grid), dim3(ATL_THREADS), (size_t)(0), (cudaStream_t)(st), grad_out, pred, label, lse, scale, C, HW, grad_pred);
    else if (C == 19) css_launch(atl_backward_kernel<19>, dim3(grid), dim3(ATL_THREADS), (size_t)(0), (cudaStream_t)(st), grad_out, pred, label, lse, scale, C, HW, grad_pred);
    else css_launch(atl_backward_kernel<0>, dim3(grid), dim3(ATL_THREADS), (size_t)(0), (cudaStream_t)(st), grad_out, pred, label, lse, scale, C, HW, grad_pred);
    CSS_CHECK_LAUNCH("css_atl_backward", 1);
    return 0;
}
