// css_grad_scatter: the backward of the contrastive loss w.r.t. the representation map.
//   grad_rep = 0 ; grad_rep[b, :, y, x] += grad_out * grad_anchor[kq, :] at every anchor pixel.
// Autograd needs the dense [B2, 256, h, w] tensor, so the compulsory traffic is one streaming write of it.
#include "css_common.cuh"

// -------------------------------------------------------------------------------------------------------------------
// Slab path.  CTA (g, b) owns every G-th 32 KB slab of image b's gradient (the image is one contiguous, 16-byte aligned
// run of 256*h*w floats).  It first lists the anchors that lie in image b (one scan of the anchor ids, a few hundred
// hits, kept in shared memory), then builds each of its slabs in shared memory -- zero fill, shared-memory atomics for
// the listed anchors' channels that fall inside (anchor pixel s touches element d*hw + s of the image for every channel
// d) -- and writes it out once with 128-bit stores.  One pass over HBM instead of a memset plus read-modify-write of
// cold sectors, and no global atomics.
// -------------------------------------------------------------------------------------------------------------------
#define GSL_ELEMS 8192                       // floats per slab (32 KB)
#define GSL_THREADS 256
#define GSL_BATCH 24                         // anchor ids a thread has in flight while the image's list is built (V321: the whole scan is one round trip)
#define GSL_PF_ANCH 3                        // listed anchors a thread may own on the pipelined path (3 x 256 per image)
#define GSL_LIST_MAX 6144                    // anchors of one image kept in shared memory (more: the slab rescans the ids)

__device__ __forceinline__ void slab_apply(float* __restrict__ slab, const float* __restrict__ grad_anchor, float go, int a, int s, int hw,
                                           int r0, int q0, int m0, int q1, int m1) {
    const int d_min = q0 + (s < m0), d_max = q1 - (s > m1);        // channels d with r0 <= d*hw + s < r1
    const float* ga = grad_anchor + (size_t)a * CSS_D;
    for (int d = d_min; d <= d_max; d += 4) {
        float g[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) g[u] = (d + u <= d_max) ? __ldg(ga + d + u) : 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (d + u <= d_max) atomicAdd(&slab[(d + u) * hw + s - r0], go * g[u]);
    }
}

__global__ void __launch_bounds__(GSL_THREADS, 4) grad_slab_kernel(const float* __restrict__ grad_out, const int32_t* __restrict__ anchor_px,
                                                                   const float* __restrict__ grad_anchor, int n_anchor, int hw,
                                                                   int list_cap, float* __restrict__ grad_rep) {
    css_pdl_enter();
    extern __shared__ __align__(16) float slab[];                  // [GSL_ELEMS] | list_a[list_cap] | list_s[list_cap]
    __shared__ int s_cnt;
    int* list_a = reinterpret_cast<int*>(slab + GSL_ELEMS);
    int* list_s = list_a + list_cap;
    const int b = blockIdx.y;
    const int img = CSS_D * hw;                                    // floats per image
    const int px_lo = b * hw, px_hi = px_lo + hw;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    for (int a0 = 0; a0 < n_anchor; a0 += GSL_THREADS * GSL_BATCH) {
        int px[GSL_BATCH];
#pragma unroll
        for (int i = 0; i < GSL_BATCH; ++i) {                      // the ids of a batch are in flight together
            const int a = a0 + i * GSL_THREADS + threadIdx.x;
            px[i] = a < n_anchor ? __ldg(anchor_px + a) : -1;
        }
#pragma unroll
        for (int i = 0; i < GSL_BATCH; ++i) {
            if (px[i] >= px_lo && px[i] < px_hi) {
                const int pos = atomicAdd(&s_cnt, 1);
                if (pos < list_cap) {
                    list_a[pos] = a0 + i * GSL_THREADS + threadIdx.x;
                    list_s[pos] = px[i] - px_lo;
                }
            }
        }
    }
    __syncthreads();
    const int cnt = s_cnt;
    const float go = __ldg(grad_out);
    float4* slab4 = reinterpret_cast<float4*>(slab);
    float* out_img = grad_rep + (size_t)b * img;
    const int n_slabs = (img + GSL_ELEMS - 1) / GSL_ELEMS;
    for (int i = threadIdx.x; i < GSL_ELEMS / 4; i += GSL_THREADS) slab4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    if (cnt <= GSL_THREADS * GSL_PF_ANCH && cnt <= list_cap && 2 * hw >= GSL_ELEMS) {
        // Pipelined path (a thread owns <= GSL_PF_ANCH listed anchors; with hw >= half a slab an anchor has <= 2 channels inside a
        // slab): the gradient values of the NEXT slab are fetched while the current slab streams out, so the L2 round trip of the
        // anchor values is off the slab's critical path (it was serialised with the write-out: apply -> barrier -> write -> barrier).
        float pv[GSL_PF_ANCH][2];
        int po[GSL_PF_ANCH][2];
        auto gather = [&](int j) {
            const int r0 = j * GSL_ELEMS, r1 = min(r0 + GSL_ELEMS, img);
            const int q0 = r0 / hw, m0 = r0 - q0 * hw, q1 = (r1 - 1) / hw, m1 = (r1 - 1) - q1 * hw;
#pragma unroll
            for (int u = 0; u < GSL_PF_ANCH; ++u) {
                const int t = threadIdx.x + u * GSL_THREADS;
                po[u][0] = po[u][1] = -1;
                if (t < cnt) {
                    const int a = list_a[t], sp = list_s[t];
                    const int d_min = q0 + (sp < m0), d_max = q1 - (sp > m1);
                    const float* ga = grad_anchor + (size_t)a * CSS_D;
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        if (d_min + c <= d_max) {
                            pv[u][c] = __ldg(ga + d_min + c);
                            po[u][c] = (d_min + c) * hw + sp - r0;
                        }
                    }
                }
            }
        };
        int j = blockIdx.x;
        if (j < n_slabs) gather(j);
        for (; j < n_slabs; j += gridDim.x) {
#pragma unroll
            for (int u = 0; u < GSL_PF_ANCH; ++u)
#pragma unroll
                for (int c = 0; c < 2; ++c)
                    if (po[u][c] >= 0) atomicAdd(&slab[po[u][c]], go * pv[u][c]);
            __syncthreads();
            const int r0 = j * GSL_ELEMS, r1 = min(r0 + GSL_ELEMS, img);
            if (j + (int)gridDim.x < n_slabs) gather(j + gridDim.x);          // in flight during the write-out below
            float4* out4 = reinterpret_cast<float4*>(out_img + r0);
            const int n4 = (r1 - r0) >> 2;
            for (int i = threadIdx.x; i < GSL_ELEMS / 4; i += GSL_THREADS) {
                if (i < n4) __stcs(out4 + i, slab4[i]);
                slab4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            __syncthreads();
        }
        return;
    }
    for (int j = blockIdx.x; j < n_slabs; j += gridDim.x) {
        const int r0 = j * GSL_ELEMS, r1 = min(r0 + GSL_ELEMS, img);
        const int q0 = r0 / hw, m0 = r0 - q0 * hw, q1 = (r1 - 1) / hw, m1 = (r1 - 1) - q1 * hw;
        if (cnt <= list_cap) {
            for (int t = threadIdx.x; t < cnt; t += GSL_THREADS) slab_apply(slab, grad_anchor, go, list_a[t], list_s[t], hw, r0, q0, m0, q1, m1);
        } else {                                                   // list overflow: rescan the ids for this slab
            for (int a = threadIdx.x; a < n_anchor; a += GSL_THREADS) {
                const int px = __ldg(anchor_px + a);
                if (px >= px_lo && px < px_hi) slab_apply(slab, grad_anchor, go, a, px - px_lo, hw, r0, q0, m0, q1, m1);
            }
        }
        __syncthreads();
        float4* out4 = reinterpret_cast<float4*>(out_img + r0);    // image start and r0 are multiples of 4 floats
        const int n4 = (r1 - r0) >> 2;                             // img is a multiple of 4, so is every slab length
        for (int i = threadIdx.x; i < GSL_ELEMS / 4; i += GSL_THREADS) {
            if (i < n4) __stcs(out4 + i, slab4[i]);
            slab4[i] = make_float4(0.f, 0.f, 0.f, 0.f);            // ready for the next slab
        }
        __syncthreads();
    }
}

// -------------------------------------------------------------------------------------------------------------------
// Fallback (very many anchors, or a gradient buffer that is not 16-byte aligned): a linear memset followed by
// <= V*Q*256 red.global.add.f32.
// -------------------------------------------------------------------------------------------------------------------
#define GS_PER_BLOCK 8
// each CTA handles GS_PER_BLOCK anchors: all pixel ids and gradient rows are loaded first (independent loads), then the
// reductions are issued; thread d owns channel d (stride h*w in the NCHW gradient)
__global__ void __launch_bounds__(CSS_D) grad_scatter_kernel(const float* __restrict__ grad_out, const int32_t* __restrict__ anchor_px,
                                                             const float* __restrict__ grad_anchor, int n_anchor, int hw,
                                                             float* __restrict__ grad_rep) {
    css_pdl_enter();
    const int base = blockIdx.x * GS_PER_BLOCK;
    int px[GS_PER_BLOCK];
    float g[GS_PER_BLOCK];
#pragma unroll
    for (int i = 0; i < GS_PER_BLOCK; ++i) px[i] = (base + i < n_anchor) ? __ldg(anchor_px + base + i) : -1;
    const float go = __ldg(grad_out);
#pragma unroll
    for (int i = 0; i < GS_PER_BLOCK; ++i)
        g[i] = (px[i] >= 0) ? ldg_stream(grad_anchor + (size_t)(base + i) * CSS_D + threadIdx.x) : 0.f;
#pragma unroll
    for (int i = 0; i < GS_PER_BLOCK; ++i) {
        if (px[i] < 0) continue;
        const int b = px[i] / hw, s = px[i] - b * hw;
        atomicAdd(grad_rep + ((size_t)b * CSS_D + threadIdx.x) * hw + s, go * g[i]);
    }
}

#define GSL_MAX_ANCHORS 32768               // every slab CTA scans the whole anchor list: beyond this the fallback is cheaper

extern "C" int css_grad_scatter(const float* grad_out, const int32_t* anchor_px, const float* grad_anchor, int n_anchor, int B2,
                                int D, int h, int w, float* grad_rep, void* stream) {
    CSS_CHECK_ARG(grad_out && anchor_px && grad_anchor && grad_rep, CSS_E_ARG, "css_grad_scatter: null pointer");
    CSS_CHECK_ARG(n_anchor > 0 && B2 > 0 && h > 0 && w > 0, CSS_E_ARG, "css_grad_scatter: non-positive size");
    CSS_CHECK_ARG(D == CSS_D, CSS_E_DIM, "css_grad_scatter: D must be %d", CSS_D);
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = (long long)B2 * D * h * w;
    if (n_anchor <= GSL_MAX_ANCHORS && ((uintptr_t)grad_rep & 15) == 0 && B2 <= 65535 && (long long)D * h * w < (1ll << 30)) {
        // list capacity: 4x the expected anchors per image (all of them when that is small); overflow only costs time
        int list_cap = 4 * ((n_anchor + B2 - 1) / B2);
        if (list_cap < 1024) list_cap = 1024;
        if (list_cap > n_anchor) list_cap = n_anchor;
        if (list_cap > GSL_LIST_MAX) list_cap = GSL_LIST_MAX;
        const int smem = GSL_ELEMS * (int)sizeof(float) + 2 * list_cap * (int)sizeof(int);
        cudaError_t e = cudaFuncSetAttribute(grad_slab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) { css_set_error("css_grad_scatter: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
        const int n_slabs = (int)(((long long)D * h * w + GSL_ELEMS - 1) / GSL_ELEMS);
        int per_sm = (227 * 1024) / (smem + 1024);
        if (per_sm > 4) per_sm = 4;
        int G = per_sm * css_sm_count() / B2;                       // all CTAs resident at once: no tail wave
        if (G > n_slabs) G = n_slabs;
        if (G < 1) G = 1;
        css_launch(grad_slab_kernel, dim3(dim3(G, B2)), dim3(GSL_THREADS), (size_t)(smem), (cudaStream_t)(st), grad_out, anchor_px, grad_anchor, n_anchor, h * w, list_cap, grad_rep);
        CSS_CHECK_LAUNCH("css_grad_scatter", 1);
        return 0;
    }
    cudaError_t e = cudaMemsetAsync(grad_rep, 0, sizeof(float) * (size_t)total, st);
    if (e != cudaSuccess) { css_set_error("css_grad_scatter: memset: %s", cudaGetErrorString(e)); return (int)e; }
    css_launch(grad_scatter_kernel, dim3((n_anchor + GS_PER_BLOCK - 1) / GS_PER_BLOCK), dim3(CSS_D), (size_t)(0), (cudaStream_t)(st), grad_out, anchor_px, grad_anchor, n_anchor, h * w,
                                                                                      grad_rep);
    CSS_CHECK_LAUNCH("css_grad_scatter", 1);
    return 0;
}


// -------------------------------------------------------------------------------------------------------------------
// Channels-last gradient: grad_rep's memory is [pixel][256], so the dense gradient is a linear zero fill plus one 1 KB row
// update per anchor (duplicates accumulate with red.global.add).
// -------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CSS_D) grad_rows_kernel(const float* __restrict__ grad_out, const int32_t* __restrict__ anchor_px,
                                                          const float* __restrict__ grad_anchor, int n_anchor, float* __restrict__ grad_rows) {
    css_pdl_enter();
    const int base = blockIdx.x * GS_PER_BLOCK;
    int px[GS_PER_BLOCK];
    float g[GS_PER_BLOCK];
#pragma unroll
    for (int i = 0; i < GS_PER_BLOCK; ++i) px[i] = (base + i < n_anchor) ? __ldg(anchor_px + base + i) : -1;
    const float go = __ldg(grad_out);
#pragma unroll
    for (int i = 0; i < GS_PER_BLOCK; ++i)
        g[i] = (px[i] >= 0) ? ldg_stream(grad_anchor + (size_t)(base + i) * CSS_D + threadIdx.x) : 0.f;
#pragma unroll
    for (int i = 0; i < GS_PER_BLOCK; ++i)
        if (px[i] >= 0) atomicAdd(grad_rows + (size_t)px[i] * CSS_D + threadIdx.x, go * g[i]);
}

extern "C" int css_grad_scatter_nhwc(const float* grad_out, const int32_t* anchor_px, const float* grad_anchor, int n_anchor, int B2, int D,
                                     int h, int w, float* grad_rows, void* stream) {
    CSS_CHECK_ARG(grad_out && anchor_px && grad_anchor && grad_rows, CSS_E_ARG, "css_grad_scatter_nhwc: null pointer");
    CSS_CHECK_ARG(n_anchor > 0 && B2 > 0 && h > 0 && w > 0, CSS_E_ARG, "css_grad_scatter_nhwc: non-positive size");
    CSS_CHECK_ARG(D == CSS_D, CSS_E_DIM, "css_grad_scatter_nhwc: D must be %d", CSS_D);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(grad_rows, 0, sizeof(float) * (size_t)B2 * D * h * w, st);
    if (e != cudaSuccess) { css_set_error("css_grad_scatter_nhwc: memset: %s", cudaGetErrorString(e)); return (int)e; }
    css_launch(grad_rows_kernel, dim3((n_anchor + GS_PER_BLOCK - 1) / GS_PER_BLOCK), dim3(CSS_D), (size_t)0, st, grad_out, anchor_px, grad_anchor, n_anchor,
               grad_rows);
    CSS_CHECK_LAUNCH("css_grad_scatter_nhwc", 1);
    return 0;
}
