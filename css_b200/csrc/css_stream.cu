// Stage 4a: per-class feature sums and counts of this rank from the pixel-major copy of the representation map: the
// all-reduce payload that replaces the reference's 116 MB all_gather (generalframeworks/loss/loss.py:77,81,102).
// HBM/L2-bound: algorithmic bytes / pixel = D*4 + 4 read (the rows were just written by the rep pass and are partly L2
// resident).
//
// Layout: css_class_blocks(N) CTAs of 64 threads, each owning a contiguous run of pixels; thread t owns channels
// 4t..4t+3 (one 128-bit load per row, 1 KB per CTA per pixel, 8 rows in flight).  Class sums accumulate in registers
// along runs of equal class sets (segmentation maps are blocky) and spill to the CTA's own [C][256] slice of `partials`
// only when the set changes; a second kernel adds the slices of the CTAs that touched a class in CTA order: no atomics,
// bit-reproducible.
#include "css_common.cuh"

#define CS_THREADS 64
#define CS_PIX 64            // pixels per CTA (upper bound of a CTA's pixel run when the grid is not capped)

extern "C" int css_class_blocks(int N) {
    const int want = (N + CS_PIX - 1) / CS_PIX;
    const int cap = css_cached_sm_count() * 32;
    return want < cap ? want : cap;
}

__device__ __forceinline__ void flush_run(float4* part, uint32_t bits, uint32_t& seen, const float4& run) {
    while (bits) {
        const int c = __ffs(bits) - 1;
        bits &= bits - 1;
        float4* dst = part + (size_t)c * (CSS_D / 4);
        if ((seen >> c) & 1u) {
            float4 o = *dst;
            o.x += run.x; o.y += run.y; o.z += run.z; o.w += run.w;
            *dst = o;
        } else {
            *dst = run;
        }
    }
}

template <typename RT>
__global__ void __launch_bounds__(CS_THREADS) class_sums_kernel(const RT* __restrict__ rows, const uint32_t* __restrict__ valid_bits,
                                                                int C, int N, float4* __restrict__ partials,
                                                                uint32_t* __restrict__ touched) {
    css_pdl_enter();
    const int t = threadIdx.x;
    const int p_begin = (int)(((long long)N * blockIdx.x) / gridDim.x);
    const int p_end = (int)(((long long)N * (blockIdx.x + 1)) / gridDim.x);
    float4* part = partials + (size_t)blockIdx.x * C * (CSS_D / 4) + t;
    float4 run = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t run_bits = 0, seen = 0;
    __shared__ uint32_t sb[CS_PIX];
    for (int c0 = p_begin; c0 < p_end; c0 += CS_PIX) {
        // the class sets of the next 64 pixels with one coalesced load; then 8 rows in flight per thread (the former
        // bits -> row dependency per 4 pixels cost two chained round trips sixteen times per CTA)
        const int len = min(CS_PIX, p_end - c0);
        __syncthreads();
        if (t < len) sb[t] = __ldg(valid_bits + c0 + t);
        __syncthreads();
        for (int i0 = 0; i0 < len; i0 += 8) {
            float4 v[8];
            uint32_t b[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                b[i] = (i0 + i < len) ? sb[i0 + i] : 0u;
                v[i] = b[i] ? row_f4(rows, (size_t)(c0 + i0 + i), t) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (b[i] != run_bits) {                          // CTA-uniform branch
                    const uint32_t rb = run_bits;
                    flush_run(part, rb, seen, run);
                    seen |= rb;
                    run = make_float4(0.f, 0.f, 0.f, 0.f);
                    run_bits = b[i];
                }
                run.x += v[i].x; run.y += v[i].y; run.z += v[i].z; run.w += v[i].w;
            }
        }
    }
    {
        const uint32_t rb = run_bits;
        flush_run(part, rb, seen, run);
        seen |= rb;
    }
    if (t == 0) touched[blockIdx.x] = seen;
}

// deterministic second stage: block (c, chunk) first compacts the ids of the CTAs that touched class c (ballot order =
// CTA order); the list is cut into CR_SEG contiguous segments, thread (seg, d) adds its segment's partials 8 loads at a
// time, and the segment sums are combined in segment order: always the same association, whatever the timing.
#define CR_SEG 16
#define CR_DCH 64                      // channels per block
__global__ void __launch_bounds__(CR_DCH * CR_SEG) class_reduce_kernel(const float* __restrict__ partials, const uint32_t* __restrict__ touched,
                                                                       const int32_t* __restrict__ meta, int G, int C,
                                                                       float* __restrict__ class_stats) {
    css_pdl_enter();
    extern __shared__ int glist[];                 // [G]
    __shared__ int wtot[CR_DCH * CR_SEG / 32];
    __shared__ float segsum[CR_SEG][CR_DCH];
    __shared__ int n_list;
    constexpr int NT = CR_DCH * CR_SEG, NW = NT / 32;
    const int c = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) n_list = 0;
    __syncthreads();
    for (int base = 0; base < G; base += NT) {
        const int g = base + tid;
        const bool on = g < G && ((touched[g] >> c) & 1u);
        const uint32_t bal = __ballot_sync(0xffffffffu, on);
        if (lane == 0) wtot[warp] = __popc(bal);
        __syncthreads();
        int off = n_list;
        for (int w2 = 0; w2 < warp; ++w2) off += wtot[w2];
        if (on) glist[off + __popc(bal & ((1u << lane) - 1u))] = g;
        __syncthreads();
        if (tid == 0) {
            int tt = 0;
            for (int w2 = 0; w2 < NW; ++w2) tt += wtot[w2];
            n_list += tt;
        }
        __syncthreads();
    }
    const int n = n_list;
    const int seg = tid / CR_DCH, dl = tid - seg * CR_DCH, d = blockIdx.y * CR_DCH + dl;
    const int per = (n + CR_SEG - 1) / CR_SEG;
    const int i0 = min(seg * per, n), i1 = min(i0 + per, n);
    float acc = 0.f;
    int i = i0;
    for (; i < i1; i += 8) {                       // predicated batches: a scalar tail would pay one L2 round trip per element
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (i + u < i1) ? partials[((size_t)glist[i + u] * C + c) * CSS_D + d] : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u];
    }
    segsum[seg][dl] = acc;
    __syncthreads();
    if (seg == 0) {
        float tot = segsum[0][dl];
#pragma unroll
        for (int s2 = 1; s2 < CR_SEG; ++s2) tot += segsum[s2][dl];
        class_stats[c * (CSS_D + 1) + d] = tot;
        if (d == 0) class_stats[c * (CSS_D + 1) + CSS_D] = (float)meta[CSS_META_N_VALID + c];
    }
}

extern "C" int css_class_stats(const void* rows, int rows_dtype, const uint32_t* valid_bits, const int32_t* meta, int N, int C, int D,
                               float* partials, uint32_t* touched, float* class_stats, void* stream) {
    CSS_CHECK_ARG(rows && valid_bits && meta && partials && touched && class_stats, CSS_E_ARG, "css_class_stats: null pointer");
    CSS_CHECK_ARG(N > 0, CSS_E_ARG, "css_class_stats: non-positive size");
    if (int e = css_check_dims(C, D)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    const int G = css_class_blocks(N);
    CSS_CHECK_ARG(rows_dtype == CSS_DTYPE_F32 || rows_dtype == CSS_DTYPE_BF16, CSS_E_DTYPE, "css_class_stats: rows dtype %d", rows_dtype);
    if (rows_dtype == CSS_DTYPE_F32)
        css_launch(class_sums_kernel<float>, dim3(G), dim3(CS_THREADS), (size_t)(0), (cudaStream_t)(st), (const float*)rows, valid_bits, C, N, (float4*)partials, touched);
    else
        css_launch(class_sums_kernel<__nv_bfloat16>, dim3(G), dim3(CS_THREADS), (size_t)(0), (cudaStream_t)(st), (const __nv_bfloat16*)rows, valid_bits, C, N, (float4*)partials, touched);
    css_launch(class_reduce_kernel, dim3(dim3(C, CSS_D / CR_DCH)), dim3(CR_DCH * CR_SEG), (size_t)(G * sizeof(int)), (cudaStream_t)(st), partials, touched, meta, G, C, class_stats);
    CSS_CHECK_LAUNCH("css_class_stats", 2);
    return 0;
}
