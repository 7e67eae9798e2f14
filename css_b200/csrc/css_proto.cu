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

__global__ void __launch_bounds__(CSS_D) proto_ema_kernel(float* __restrict__ protos, const float* __restrict__ stats,
                                                          const int32_t* __restrict__ meta, float alpha, float one_minus_alpha,
                                                          float temp, int C, float* __restrict__ proto_hat, float* __restrict__ cdf) {
    __shared__ float ph[CSS_CMAX * CSS_D];          // normalised updated prototypes
    __shared__ float sim[CSS_CMAX * CSS_CMAX];
    __shared__ float scratch[8];
    const int d = threadIdx.x;
    for (int c = 0; c < C; ++c) {
        float p = protos[c * CSS_D + d];
        if (meta[CSS_META_N_VALID + c] > 0) {                                   // only classes present on THIS rank (loss.py:96-97)
            const float rowsum = block_sum_256(p, scratch);
            const float mean = __fdiv_rn(stats[c * (CSS_D + 1) + d], stats[c * (CSS_D + 1) + CSS_D]);   // loss.py:102
            p = (rowsum == 0.f) ? mean                                           // first touch (loss.py:103-105)
                                : __fadd_rn(__fmul_rn(alpha, p), __fmul_rn(one_minus_alpha, mean));   // loss.py:108
            protos[c * CSS_D + d] = p;
        }
        const float n2 = block_sum_256(p * p, scratch);
        const float ph_cd = __fdiv_rn(p, fmaxf(sqrtf(n2), 1e-8f));              // cosine_similarity eps (loss.py:134,146)
        ph[c * CSS_D + d] = ph_cd;
        proto_hat[c * CSS_D + d] = ph_cd;
    }
    __syncthreads();
    const int V = meta[CSS_META_V];
    const int warp = d >> 5, lane = d & 31;
    // cos(P_k, P_j) for every ordered pair of present classes, j in rotated order k+1..V-1,0..k-1
    for (int pair = warp; pair < V * (V - 1); pair += 8) {
        const int k = pair / (V - 1), i = pair - k * (V - 1);
        const int ck = meta[CSS_META_CLS_OF_SLOT + k];
        const int cj = meta[CSS_META_CLS_OF_SLOT + (k + 1 + i) % V];
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < CSS_D / 32; ++e) s = fmaf(ph[ck * CSS_D + lane + 32 * e], ph[cj * CSS_D + lane + 32 * e], s);
        s = warp_sum(s);
        if (lane == 0) sim[k * CSS_CMAX + i] = s;
    }
    __syncthreads();
    if (d < CSS_CMAX) {
        const int k = d;
        if (k < V && V > 1) {
            float m = -INFINITY;
            for (int i = 0; i < V - 1; ++i) {
                sim[k * CSS_CMAX + i] = __fdiv_rn(sim[k * CSS_CMAX + i], temp);
                m = fmaxf(m, sim[k * CSS_CMAX + i]);
            }
            float tot = 0.f;
            for (int i = 0; i < V - 1; ++i) {
                sim[k * CSS_CMAX + i] = expf(sim[k * CSS_CMAX + i] - m);
                tot += sim[k * CSS_CMAX + i];
            }
            float run = 0.f;
            for (int i = 0; i < CSS_CMAX; ++i) {
                if (i < V - 1) run += __fdiv_rn(sim[k * CSS_CMAX + i], tot);
                cdf[k * CSS_CMAX + i] = (i < V - 2) ? run : 1.f;               // last bin absorbs rounding
            }
        } else {
            for (int i = 0; i < CSS_CMAX; ++i) cdf[k * CSS_CMAX + i] = 1.f;
        }
    }
}

extern "C" int css_proto_ema(float* prototypes, const float* class_stats, const int32_t* meta, float alpha, float one_minus_alpha,
                             float temp, int C, int D, float* proto_hat, float* class_cdf, void* stream) {
    CSS_CHECK_ARG(prototypes && class_stats && meta && proto_hat && class_cdf, CSS_E_ARG, "css_proto_ema: null pointer");
    if (int e = css_check_dims(C, D)) return e;
    proto_ema_kernel<<<1, CSS_D, 0, (cudaStream_t)stream>>>(prototypes, class_stats, meta, alpha, one_minus_alpha, temp, C,
                                                           proto_hat, class_cdf);
    CSS_CHECK_LAUNCH("css_proto_ema", 1);
    return 0;
}
