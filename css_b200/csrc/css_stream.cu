// Stage 4a: ONE streaming read of the NCHW representation map that produces
//   (1) the per-class feature sums of this rank (all-reduce payload replacing the reference's 116 MB all_gather,
//       generalframeworks/loss/loss.py:77,81,102),
//   (2) ||x_p|| per pixel and the pixel-major copy x_p / max(||x_p||, 1e-8) that the scoring kernel gathers 1 KB rows
//       from (loss.py:85,111-112,142,146: permute + boolean gathers + cosine_similarity's normalisation).
// HBM-bound: algorithmic bytes / pixel = D*4 read + D*4 written + 4 (norm) + 4 (class set).
//
// Layout: a persistent grid of css_stream_blocks() CTAs, each owning a contiguous run of 32-pixel tiles.  A tile is
// staged through shared memory ([256 ch][33]) by coalesced 128 B channel-row reads; the transposed write-out gives every
// thread one channel column, so class sums accumulate in registers along runs of equal class sets (segmentation maps are
// blocky) and spill to a per-CTA [C][256] shared accumulator only when the set changes: no atomics, deterministic.
#include "css_common.cuh"

#define ST_PIX 32
#define ST_PAD 33
#define ST_THREADS 256

extern "C" int css_stream_blocks(void) { return css_cached_sm_count() * 4; }

__device__ __forceinline__ void flush_run(float* sums, uint32_t bits, float run, int d) {
    while (bits) {
        const int c = __ffs(bits) - 1;
        bits &= bits - 1;
        sums[c * CSS_D + d] += run;
    }
}

__global__ void __launch_bounds__(ST_THREADS) stream_rep_kernel(const float* __restrict__ rep, const uint32_t* __restrict__ valid_bits,
                                                                int C, int hw, int N, float* __restrict__ rows_hat,
                                                                float* __restrict__ norms, float* __restrict__ partials,
                                                                uint32_t* __restrict__ touched) {
    extern __shared__ float smem[];
    float* tile = smem;                                   // [256][33]
    float* red = tile + CSS_D * ST_PAD;                   // [8][32] partial squared norms
    float* inv = red + 8 * ST_PIX;                        // [32]
    uint32_t* bits_s = reinterpret_cast<uint32_t*>(inv + ST_PIX);   // [32]
    float* sums = reinterpret_cast<float*>(bits_s + ST_PIX);        // [C][256]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < C * CSS_D; i += ST_THREADS) sums[i] = 0.f;

    const int n_tiles = (N + ST_PIX - 1) / ST_PIX;
    const int t_begin = (int)(((long long)n_tiles * blockIdx.x) / gridDim.x);
    const int t_end = (int)(((long long)n_tiles * (blockIdx.x + 1)) / gridDim.x);

    float run = 0.f;
    uint32_t run_bits = 0, seen = 0;
    for (int t = t_begin; t < t_end; ++t) {
        const int p0 = t * ST_PIX;
        // ---- load: warp w reads channels [32w, 32w+32), lane = pixel; 32 independent 128 B rows in flight per warp
        {
            const int p = p0 + lane;
            const bool ok = p < N;
            const int b = ok ? p / hw : 0, s = ok ? p - b * hw : 0;
            const float* x = rep + ((size_t)b * CSS_D + warp * 32) * hw + s;
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = ok ? ldg_stream(x + (size_t)i * hw) : 0.f;
            float n2 = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                tile[(warp * 32 + i) * ST_PAD + lane] = v[i];
                n2 = fmaf(v[i], v[i], n2);
            }
            red[warp * ST_PIX + lane] = n2;
            if (warp == 0) bits_s[lane] = ok ? valid_bits[p] : 0u;
        }
        __syncthreads();
        if (warp == 0) {
            float n2 = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) n2 += red[k * ST_PIX + lane];
            const float nrm = sqrtf(n2);
            inv[lane] = 1.f / fmaxf(nrm, 1e-8f);           // cosine_similarity eps (loss.py:146)
            if (p0 + lane < N) norms[p0 + lane] = nrm;
        }
        __syncthreads();
        // ---- transposed write-out + class sums: thread = channel d, loop over the tile's pixels
        {
            const int d = tid;
            const int n_here = min(ST_PIX, N - p0);
            float* out = rows_hat + (size_t)p0 * CSS_D + d;
#pragma unroll 8
            for (int i = 0; i < n_here; ++i) {
                const float v = tile[d * ST_PAD + i];
                out[(size_t)i * CSS_D] = v * inv[i];
                const uint32_t bits = bits_s[i];
                if (bits != run_bits) {                     // block-uniform branch
                    flush_run(sums, run_bits, run, d);
                    seen |= run_bits;
                    run = 0.f;
                    run_bits = bits;
                }
                run += v;
            }
        }
        __syncthreads();
    }
    flush_run(sums, run_bits, run, tid);
    seen |= run_bits;
    // per-CTA partial sums of the classes this CTA touched (thread d only ever touched sums[*][d]: no sync needed)
    for (uint32_t u = seen; u;) {
        const int c = __ffs(u) - 1;
        u &= u - 1;
        partials[((size_t)blockIdx.x * C + c) * CSS_D + tid] = sums[c * CSS_D + tid];
    }
    if (tid == 0) touched[blockIdx.x] = seen;
}

// deterministic second stage: block c first compacts the ids of the CTAs that touched class c (ballot order = CTA order),
// then thread d adds their partials 8 loads at a time, always in the same order.
__global__ void __launch_bounds__(CSS_D) stream_reduce_kernel(const float* __restrict__ partials, const uint32_t* __restrict__ touched,
                                                              const int32_t* __restrict__ meta, int G, int C,
                                                              float* __restrict__ class_stats) {
    extern __shared__ int glist[];                 // [G]
    __shared__ int wtot[CSS_D / 32];
    __shared__ int n_list;
    const int c = blockIdx.x, d = threadIdx.x, warp = d >> 5, lane = d & 31;
    if (d == 0) n_list = 0;
    __syncthreads();
    for (int base = 0; base < G; base += CSS_D) {
        const int g = base + d;
        const bool on = g < G && ((touched[g] >> c) & 1u);
        const uint32_t bal = __ballot_sync(0xffffffffu, on);
        if (lane == 0) wtot[warp] = __popc(bal);
        __syncthreads();
        int off = n_list;
        for (int w2 = 0; w2 < warp; ++w2) off += wtot[w2];
        if (on) glist[off + __popc(bal & ((1u << lane) - 1u))] = g;
        __syncthreads();
        if (d == 0) {
            int t = 0;
            for (int w2 = 0; w2 < CSS_D / 32; ++w2) t += wtot[w2];
            n_list += t;
        }
        __syncthreads();
    }
    const int n = n_list;
    float acc = 0.f;
    int i = 0;
    for (; i + 8 <= n; i += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = partials[((size_t)glist[i + u] * C + c) * CSS_D + d];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u];
    }
    for (; i < n; ++i) acc += partials[((size_t)glist[i] * C + c) * CSS_D + d];
    class_stats[c * (CSS_D + 1) + d] = acc;
    if (d == 0) class_stats[c * (CSS_D + 1) + CSS_D] = (float)meta[CSS_META_N_VALID + c];
}

extern "C" int css_stream_rep(const void* rep, int rep_dtype, const uint32_t* valid_bits, const int32_t* meta, int B2, int C,
                              int D, int h, int w, float* rows_hat, float* norms, float* partials, uint32_t* touched,
                              float* class_stats, void* stream) {
    CSS_CHECK_ARG(rep && valid_bits && meta && rows_hat && norms && partials && touched && class_stats, CSS_E_ARG,
                  "css_stream_rep: null pointer");
    CSS_CHECK_ARG(B2 > 0 && h > 0 && w > 0, CSS_E_ARG, "css_stream_rep: non-positive size");
    if (int e = css_check_dims(C, D)) return e;
    CSS_CHECK_ARG(rep_dtype == CSS_DTYPE_F32, CSS_E_DTYPE, "css_stream_rep: rep dtype %d not supported", rep_dtype);
    CSS_CHECK_ARG((long long)B2 * h * w * CSS_CMAX < (1ll << 31), CSS_E_SIZE, "css_stream_rep: too many pixels");
    cudaStream_t st = (cudaStream_t)stream;
    const int hw = h * w, N = B2 * hw, G = css_stream_blocks();
    const size_t smem = (size_t)(CSS_D * ST_PAD + 8 * ST_PIX + 2 * ST_PIX + C * CSS_D) * sizeof(float);
    {   // > 48 KB of dynamic shared memory needs the opt-in (host-side attribute, legal during graph capture)
        cudaError_t e = cudaFuncSetAttribute(stream_rep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)((CSS_D * ST_PAD + 10 * ST_PIX + CSS_CMAX * CSS_D) * sizeof(float)));
        if (e != cudaSuccess) { css_set_error("css_stream_rep: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    }
    stream_rep_kernel<<<G, ST_THREADS, smem, st>>>((const float*)rep, valid_bits, C, hw, N, rows_hat, norms, partials, touched);
    stream_reduce_kernel<<<C, CSS_D, G * sizeof(int), st>>>(partials, touched, meta, G, C, class_stats);
    CSS_CHECK_LAUNCH("css_stream_rep", 2);
    return 0;
}
