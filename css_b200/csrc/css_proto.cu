// Stage 4b: prototype EMA update (alpha_p) from the all-reduced per-class sums / counts, and the tables the scoring
// kernel needs (normalised prototypes, class-sampling CDFs).
// Reference: generalframeworks/loss/loss.py:101-109 (mean / first-touch / EMA, in place), :133-135 (proto_prob).
#include "css_common.cuh"

__device__ __forceinline__ float block_sum_256(float v, float* scratch) {
    // fixed-order tree: warp shuffles then 8 partials -> deterministic
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += scratch[i];
    return s;
}

// one CTA per class: mean / first-touch / EMA in place, then the normalised row for the scorer
__global__ void __launch_bounds__(CSS_D) proto_ema_kernel(float* __restrict__ protos, const float* __restrict__ stats,
                                                          const int32_t* __restrict__ meta, float alpha, float one_minus_alpha,
                                                          int update_rule, float* __restrict__ proto_hat) {
    css_pdl_enter();
    __shared__ float scratch[8];
    const int c = blockIdx.x, d = threadIdx.x;
    float p = protos[c * CSS_D + d];
    const float s_d = stats[c * (CSS_D + 1) + d], s_n = stats[c * (CSS_D + 1) + CSS_D];    // in flight with p and meta, ahead of the barriers
    // CSS_UPDATE_LOCAL: only classes present on THIS rank (loss.py:96-97); CSS_UPDATE_GLOBAL: every class any rank saw, so
    // that ranks starting from equal prototypes stay bit-identical (the summed statistics are the same on every rank)
    const bool touch = update_rule == CSS_UPDATE_GLOBAL ? (s_n > 0.f) : (meta[CSS_META_N_VALID + c] > 0);
    if (touch) {
        const float rowsum = block_sum_256(p, scratch);
        const float mean = __fdiv_rn(s_d, s_n);                              // loss.py:102
        p = (rowsum == 0.f) ? mean                                           // first touch (loss.py:103-105)
                            : __fadd_rn(__fmul_rn(alpha, p), __fmul_rn(one_minus_alpha, mean));   // loss.py:108
        protos[c * CSS_D + d] = p;
    }
    const float n2 = block_sum_256(p * p, scratch);
    proto_hat[c * CSS_D + d] = __fdiv_rn(p, fmaxf(sqrtf(n2), 1e-8f));       // cosine_similarity eps (loss.py:134,146)
}

// one CTA per present-class slot k: cos(P_k, P_j) for the other present classes j in rotated order k+1..V-1,0..k-1,
// softmax(/temp) and its inclusive CDF (loss.py:133-135)
__global__ void __launch_bounds__(CSS_D) class_cdf_kernel(const float* __restrict__ proto_hat, const int32_t* __restrict__ meta,
                                                          float temp, float* __restrict__ cdf) {
    css_pdl_enter();
    __shared__ float sim[CSS_CMAX];
    const int k = blockIdx.x, d = threadIdx.x, warp = d >> 5, lane = d & 31;
    const int V = meta[CSS_META_V];
    if (k >= V || V <= 1) {
        if (d < CSS_CMAX) cdf[k * CSS_CMAX + d] = 1.f;
        return;
    }
    const int ck = meta[CSS_META_CLS_OF_SLOT + k];
    // warp w scores the other classes i = w, w + 8, w + 16, w + 24 (V - 1 <= 31): their class ids first, then all their rows, so
    // the warp pays one dependent pair of L2 round trips instead of one pair per class
    constexpr int NW = CSS_D / 32, PER = CSS_CMAX / NW;
    int cj[PER];
#pragma unroll
    for (int t = 0; t < PER; ++t) {
        const int i = warp + t * NW;
        cj[t] = (i < V - 1) ? meta[CSS_META_CLS_OF_SLOT + (k + 1 + i) % V] : ck;
    }
    float a[CSS_D / 32], b[PER][CSS_D / 32];
#pragma unroll
    for (int e = 0; e < CSS_D / 32; ++e) a[e] = proto_hat[ck * CSS_D + lane + 32 * e];
#pragma unroll
    for (int t = 0; t < PER; ++t)
#pragma unroll
        for (int e = 0; e < CSS_D / 32; ++e) b[t][e] = proto_hat[cj[t] * CSS_D + lane + 32 * e];
#pragma unroll
    for (int t = 0; t < PER; ++t) {
        const int i = warp + t * NW;
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < CSS_D / 32; ++e) s = fmaf(a[e], b[t][e], s);
        s = warp_sum(s);
        if (lane == 0 && i < V - 1) sim[i] = __fdiv_rn(s, temp);
    }
    __syncthreads();
    if (warp == 0) {
        // lane i owns class i: max, exp and the division run in parallel; the two running sums keep their sequential order
        // (lane 0), so the table is bit-identical to a single-thread evaluation
        const bool live = lane < V - 1;
        float v = live ? sim[lane] : -INFINITY, m = v;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        const float ex = live ? expf(v - m) : 0.f;
        if (live) sim[lane] = ex;
        __syncwarp();
        float tot = 0.f;
        if (lane == 0)
            for (int i = 0; i < V - 1; ++i) tot += sim[i];
        tot = __shfl_sync(0xffffffffu, tot, 0);
        const float pr = live ? __fdiv_rn(ex, tot) : 0.f;
        __syncwarp();
        sim[lane] = pr;
        __syncwarp();
        if (lane == 0) {
            float run = 0.f;
            for (int i = 0; i < CSS_CMAX; ++i) {
                if (i < V - 1) run += sim[i];
                cdf[k * CSS_CMAX + i] = (i < V - 2) ? run : 1.f;             // last bin absorbs rounding
            }
        }
    }
}

extern "C" int css_proto_ema(float* prototypes, const float* class_stats, const int32_t* meta, float alpha, float one_minus_alpha,
                             float temp, int update_rule, int C, int D, float* proto_hat, float* class_cdf, void* stream) {
    CSS_CHECK_ARG(prototypes && class_stats && meta && proto_hat && class_cdf, CSS_E_ARG, "css_proto_ema: null pointer");
    CSS_CHECK_ARG(update_rule == CSS_UPDATE_LOCAL || update_rule == CSS_UPDATE_GLOBAL, CSS_E_ARG, "css_proto_ema: bad update_rule %d", update_rule);
    if (int e = css_check_dims(C, D)) return e;
    css_launch(proto_ema_kernel, dim3(C), dim3(CSS_D), (size_t)(0), (cudaStream_t)((cudaStream_t)stream), prototypes, class_stats, meta, alpha, one_minus_alpha, update_rule, proto_hat);
    css_launch(class_cdf_kernel, dim3(CSS_CMAX), dim3(CSS_D), (size_t)(0), (cudaStream_t)((cudaStream_t)stream), proto_hat, meta, temp, class_cdf);
    CSS_CHECK_LAUNCH("css_proto_ema", 2);
    return 0;
}
