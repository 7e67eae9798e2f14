// Development harness 3: cp.async-staged variant of the similarity kernel.  Not part of the product library.
// A CTA owns a tile of TP pixels of one image and walks the 256 channel planes in chunks of CK planes through an
// S-stage shared-memory ring filled by 16-byte cp.async (planes are only 4-byte aligned, so every plane's window is
// widened to 16-byte boundaries and lands in shared memory with the same phase).  Bytes in flight live in shared
// memory instead of registers.
#include <cuda_runtime.h>
#include <stdint.h>
#define CSS_D 256
#define CSS_CMAX 32

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NG, int TP, int CK, int S, int MINB, int MODE, int KS>
__global__ void __launch_bounds__(TP / 2 * KS, MINB) sim3_kernel(const float* __restrict__ rep, const float* __restrict__ scratch, int hw,
                                                            int B, int C, long long total_elems, float* __restrict__ out) {
    constexpr int NT = TP / 2 * KS;            // threads: pixels i and i + TP/2 with i = t % (TP/2), channel slice = t / (TP/2)
    constexpr int PITCH = TP + 8;              // floats per plane in a stage (TP + <= 3 phase + pad, multiple of 4)
    constexpr int NCH = CSS_D / CK;            // chunks per tile
    constexpr int HC = CK / KS;                // channels per slice per chunk
    constexpr int TPP = NT / CK;               // threads per plane when loading
    constexpr int MAXJ = (TP / 4 + 1 + TPP - 1) / TPP;
    extern __shared__ __align__(16) float smem[];
    float4* sp = reinterpret_cast<float4*>(smem);                  // [256][NG] prototypes
    float* ring = smem + CSS_D * NG * 4;                           // [S][CK][PITCH]
    float* comb = ring + S * CK * PITCH;                           // [KS-1][TP][4*NG + 1]
    for (int i = threadIdx.x; i < CSS_D * NG; i += NT) {
        const int d = i / NG, g = i - d * NG;
        sp[i] = reinterpret_cast<const float4*>(scratch)[d * (CSS_CMAX / 4) + g];
    }
    const int tiles_per_img = (hw + TP - 1) / TP;
    const int n_tiles = tiles_per_img * B;
    const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int n_seq = my_tiles * NCH;
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
    const int lp = threadIdx.x / TPP, lj = threadIdx.x % TPP;      // loader role: plane within chunk, first 16 B column

    auto issue = [&](int seq) {
        if (MODE != 2 && seq < n_seq) {
            const int tile = blockIdx.x + (seq / NCH) * gridDim.x, kc = seq % NCH;
            const int img = tile / tiles_per_img, p0 = (tile - img * tiles_per_img) * TP;
            const int npx = min(TP, hw - p0);
            const long long a0 = ((long long)img * CSS_D + kc * CK + lp) * hw + p0;
            const int phase = (int)(a0 & 3);
            const long long base = a0 - phase;
            const int n16 = (phase + npx + 3) >> 2;
            const uint32_t dst = ring_s + (((seq % S) * CK + lp) * PITCH) * 4;
#pragma unroll
            for (int m = 0; m < MAXJ; ++m) {
                const int j = lj + m * TPP;
                if (j < n16) {
                    const long long e = base + 4 * j;
                    const long long left = total_elems - e;
                    cp_async16(dst + j * 16, rep + e, left >= 4 ? 16 : (left > 0 ? (int)left * 4 : 0));
                }
            }
        }
        cp_commit();
    };

#pragma unroll
    for (int s = 0; s < S - 1; ++s) issue(s);

    constexpr int HP = TP / 2;
    const int i = threadIdx.x % HP, hf = threadIdx.x / HP;
    const int r4 = hw & 3;
    float2 acc[2][2 * NG];
    float n2[2];
    for (int seq = 0; seq < n_seq; ++seq) {
        const int kc = seq % NCH;
        const int tile = blockIdx.x + (seq / NCH) * gridDim.x;
        const int img = tile / tiles_per_img, p0 = (tile - img * tiles_per_img) * TP;
        if (kc == 0) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                n2[j] = 0.f;
#pragma unroll
                for (int a = 0; a < 2 * NG; ++a) acc[j][a] = make_float2(0.f, 0.f);
            }
        }
        cp_wait<S - 2>();
        __syncthreads();                       // chunk `seq` has landed for everyone; everyone finished chunk seq-1
        issue(seq + S - 1);                    // refill the stage chunk seq-1 used
        // plane c of the chunk sits at phase (p0 + c * hw) & 3  (the chunk's first plane index is a multiple of 4)
        const float* st = ring + ((seq % S) * CK + hf * HC) * PITCH + i;
        int ph[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) ph[k] = (p0 + (hf * HC + k) * r4) & 3;
        const float4* pr = sp + (kc * CK + hf * HC) * NG;
#pragma unroll
        for (int u = 0; u < (MODE == 1 ? 1 : HC); ++u) {
            const float v0 = st[u * PITCH + ph[u & 3]], v1 = st[u * PITCH + ph[u & 3] + HP];
            n2[0] = fmaf(v0, v0, n2[0]);
            n2[1] = fmaf(v1, v1, n2[1]);
            const float2 vv0 = make_float2(v0, v0), vv1 = make_float2(v1, v1);
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const float4 q = pr[u * NG + g];
                acc[0][2 * g + 0] = __ffma2_rn(vv0, make_float2(q.x, q.y), acc[0][2 * g + 0]);
                acc[0][2 * g + 1] = __ffma2_rn(vv0, make_float2(q.z, q.w), acc[0][2 * g + 1]);
                acc[1][2 * g + 0] = __ffma2_rn(vv1, make_float2(q.x, q.y), acc[1][2 * g + 0]);
                acc[1][2 * g + 1] = __ffma2_rn(vv1, make_float2(q.z, q.w), acc[1][2 * g + 1]);
            }
        }
        if (kc == NCH - 1) {                   // tile done: combine the channel slices, normalise, write
            constexpr int CW = 4 * NG + 1;
            if (hf > 0) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    float* cb = comb + ((hf - 1) * TP + i + j * HP) * CW;
#pragma unroll
                    for (int a = 0; a < 2 * NG; ++a) {
                        cb[2 * a] = acc[j][a].x;
                        cb[2 * a + 1] = acc[j][a].y;
                    }
                    cb[4 * NG] = n2[j];
                }
            }
            __syncthreads();
            if (hf == 0) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int px = p0 + i + j * HP;
                    if (px >= hw) continue;
                    float nn = n2[j];
                    float2 t[2 * NG];
#pragma unroll
                    for (int a = 0; a < 2 * NG; ++a) t[a] = acc[j][a];
#pragma unroll
                    for (int s2 = 0; s2 < KS - 1; ++s2) {
                        const float* cb = comb + (s2 * TP + i + j * HP) * CW;
#pragma unroll
                        for (int a = 0; a < 2 * NG; ++a) {
                            t[a].x += cb[2 * a];
                            t[a].y += cb[2 * a + 1];
                        }
                        nn += cb[4 * NG];
                    }
                    const float inv = 1.f / fmaxf(sqrtf(nn), 1e-12f);
                    float* o = out + (size_t)img * C * hw + px;
#pragma unroll
                    for (int a = 0; a < 2 * NG; ++a) {
                        if (2 * a < C) o[(size_t)(2 * a) * hw] = t[a].x * inv;
                        if (2 * a + 1 < C) o[(size_t)(2 * a + 1) * hw] = t[a].y * inv;
                    }
                }
            }
            // comb is rewritten only after the next tile's NCH barriers
        }
    }
    cp_wait<0>();
}

template <int TP, int CK, int S, int MINB, int KS, int MODE = 0>
static int launch(const float* rep, const float* scratch, int hw, int N, int C, float* out, cudaStream_t st) {
    constexpr int NG = 6;
    const int B = N / hw;
    const size_t smem = (size_t)(CSS_D * NG * 4 + S * CK * (TP + 8) + (KS - 1) * TP * (4 * NG + 1)) * sizeof(float);
    auto k = sim3_kernel<NG, TP, CK, S, MINB, MODE, KS>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int n_tiles = ((hw + TP - 1) / TP) * B;
    int per_sm = (int)(227 * 1024 / (smem + 1024));
    if (per_sm > MINB) per_sm = MINB;
    const int grid = n_tiles < 148 * per_sm ? n_tiles : 148 * per_sm;
    k<<<grid, TP / 2 * KS, smem, st>>>(rep, scratch, hw, B, C, (long long)N * CSS_D, out);
    return 0;
}


// ---- variant family 2: prototypes as CONSTANT-bank operands of scalar FFMA (2 register reads per FMA instead of 3) ----
__constant__ float4 c_proto[CSS_D * 6];

template <int NG, int TP, int CK, int S, int KS, int HF>
__device__ __forceinline__ void chunk_const(const float* __restrict__ st, const int (&ph)[4], int kc, float (&acc)[2][4 * NG], float (&n2)[2]) {
    constexpr int HC = CK / KS, HP = TP / 2, PITCH = TP + 8;
#pragma unroll
    for (int u = 0; u < HC; ++u) {
        const float v0 = st[u * PITCH + ph[u & 3]], v1 = st[u * PITCH + ph[u & 3] + HP];
        n2[0] = fmaf(v0, v0, n2[0]);
        n2[1] = fmaf(v1, v1, n2[1]);
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const float4 q = c_proto[(kc * CK + HF * HC + u) * NG + g];
            acc[0][4 * g + 0] = fmaf(v0, q.x, acc[0][4 * g + 0]);
            acc[0][4 * g + 1] = fmaf(v0, q.y, acc[0][4 * g + 1]);
            acc[0][4 * g + 2] = fmaf(v0, q.z, acc[0][4 * g + 2]);
            acc[0][4 * g + 3] = fmaf(v0, q.w, acc[0][4 * g + 3]);
            acc[1][4 * g + 0] = fmaf(v1, q.x, acc[1][4 * g + 0]);
            acc[1][4 * g + 1] = fmaf(v1, q.y, acc[1][4 * g + 1]);
            acc[1][4 * g + 2] = fmaf(v1, q.z, acc[1][4 * g + 2]);
            acc[1][4 * g + 3] = fmaf(v1, q.w, acc[1][4 * g + 3]);
        }
    }
}

template <int NG, int TP, int CK, int S, int MINB, int MODE, int KS>
__global__ void __launch_bounds__(TP / 2 * KS, MINB) sim4_kernel(const float* __restrict__ rep, int hw, int B, int C, long long total_elems,
                                                                 float* __restrict__ out) {
    constexpr int NT = TP / 2 * KS;
    constexpr int PITCH = TP + 8;
    constexpr int NCH = CSS_D / CK;
    constexpr int HC = CK / KS;
    constexpr int TPP = NT / CK;
    constexpr int MAXJ = (TP / 4 + 1 + TPP - 1) / TPP;
    constexpr int HP = TP / 2;
    constexpr int CW = 4 * NG + 1;
    extern __shared__ __align__(16) float smem[];
    float* ring = smem;                                            // [S][CK][PITCH]
    float* comb = ring + S * CK * PITCH;                           // [KS-1][TP][CW]
    const int tiles_per_img = (hw + TP - 1) / TP;
    const int n_tiles = tiles_per_img * B;
    const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int n_seq = my_tiles * NCH;
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
    const int lp = threadIdx.x / TPP, lj = threadIdx.x % TPP;

    auto issue = [&](int seq) {
        if (MODE != 2 && seq < n_seq) {
            const int tile = blockIdx.x + (seq / NCH) * gridDim.x, kc = seq % NCH;
            const int img = tile / tiles_per_img, p0 = (tile - img * tiles_per_img) * TP;
            const int npx = min(TP, hw - p0);
            const long long a0 = ((long long)img * CSS_D + kc * CK + lp) * hw + p0;
            const int phase = (int)(a0 & 3);
            const long long base = a0 - phase;
            const int n16 = (phase + npx + 3) >> 2;
            const uint32_t dst = ring_s + (((seq % S) * CK + lp) * PITCH) * 4;
#pragma unroll
            for (int m = 0; m < MAXJ; ++m) {
                const int j = lj + m * TPP;
                if (j < n16) {
                    const long long e = base + 4 * j;
                    const long long left = total_elems - e;
                    cp_async16(dst + j * 16, rep + e, left >= 4 ? 16 : (left > 0 ? (int)left * 4 : 0));
                }
            }
        }
        cp_commit();
    };
#pragma unroll
    for (int s = 0; s < S - 1; ++s) issue(s);

    const int i = threadIdx.x % HP, hf = threadIdx.x / HP;
    const int r4 = hw & 3;
    float acc[2][4 * NG];
    float n2[2];
    for (int seq = 0; seq < n_seq; ++seq) {
        const int kc = seq % NCH;
        const int tile = blockIdx.x + (seq / NCH) * gridDim.x;
        const int img = tile / tiles_per_img, p0 = (tile - img * tiles_per_img) * TP;
        if (kc == 0) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                n2[j] = 0.f;
#pragma unroll
                for (int a = 0; a < 4 * NG; ++a) acc[j][a] = 0.f;
            }
        }
        cp_wait<S - 2>();
        __syncthreads();
        issue(seq + S - 1);
        const float* st = ring + ((seq % S) * CK + hf * HC) * PITCH + i;
        int ph[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) ph[k] = (p0 + (hf * HC + k) * r4) & 3;
        if (MODE != 1) {
            if (KS == 1 || hf == 0) chunk_const<NG, TP, CK, S, KS, 0>(st, ph, kc, acc, n2);
            else if (KS == 2 || hf == 1) chunk_const<NG, TP, CK, S, KS, (KS > 1 ? 1 : 0)>(st, ph, kc, acc, n2);
            else if (hf == 2) chunk_const<NG, TP, CK, S, KS, (KS > 2 ? 2 : 0)>(st, ph, kc, acc, n2);
            else chunk_const<NG, TP, CK, S, KS, (KS > 3 ? 3 : 0)>(st, ph, kc, acc, n2);
        } else {
            n2[0] += st[ph[0]];
        }
        if (kc == NCH - 1) {
            if (hf > 0) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    float* cb = comb + ((hf - 1) * TP + i + j * HP) * CW;
#pragma unroll
                    for (int a = 0; a < 4 * NG; ++a) cb[a] = acc[j][a];
                    cb[4 * NG] = n2[j];
                }
            }
            __syncthreads();
            if (hf == 0) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int px = p0 + i + j * HP;
                    if (px >= hw) continue;
                    float nn = n2[j];
                    float t[4 * NG];
#pragma unroll
                    for (int a = 0; a < 4 * NG; ++a) t[a] = acc[j][a];
#pragma unroll
                    for (int s2 = 0; s2 < KS - 1; ++s2) {
                        const float* cb = comb + (s2 * TP + i + j * HP) * CW;
#pragma unroll
                        for (int a = 0; a < 4 * NG; ++a) t[a] += cb[a];
                        nn += cb[4 * NG];
                    }
                    const float inv = 1.f / fmaxf(sqrtf(nn), 1e-12f);
                    float* o = out + (size_t)img * C * hw + px;
#pragma unroll
                    for (int a = 0; a < 4 * NG; ++a)
                        if (a < C) o[(size_t)a * hw] = t[a] * inv;
                }
            }
        }
    }
    cp_wait<0>();
}

template <int TP, int CK, int S, int MINB, int KS, int MODE = 0>
static int launch4(const float* rep, const float* scratch, int hw, int N, int C, float* out, cudaStream_t st) {
    constexpr int NG = 6;
    const int B = N / hw;
    // [256][32] scratch -> [256][24] constant image
    cudaError_t e = cudaMemcpy2DAsync(c_proto, NG * 16, scratch, CSS_CMAX * 4, NG * 16, CSS_D, cudaMemcpyDeviceToDevice, st);
    (void)e;
    void* sym;
    cudaGetSymbolAddress(&sym, c_proto);
    e = cudaMemcpy2DAsync(sym, NG * 16, scratch, CSS_CMAX * 4, NG * 16, CSS_D, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return (int)e;
    const size_t smem = (size_t)(S * CK * (TP + 8) + (KS > 1 ? (KS - 1) : 1) * TP * (4 * NG + 1)) * sizeof(float);
    auto k = sim4_kernel<NG, TP, CK, S, MINB, MODE, KS>;
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int n_tiles = ((hw + TP - 1) / TP) * B;
    int per_sm = (int)(227 * 1024 / (smem + 1024));
    if (per_sm > MINB) per_sm = MINB;
    const int grid = n_tiles < 148 * per_sm ? n_tiles : 148 * per_sm;
    k<<<grid, TP / 2 * KS, smem, st>>>(rep, hw, B, C, (long long)N * CSS_D, out);
    return 0;
}


// ---- variant family 3: register path (no staging), classes split across the two half-warps, 4 pixels per lane ----
// lane = ch * 16 + pl: class half ch owns 12 of the 24 class slots, so a channel costs 3 LDS.128 per 64 pixel-channels
// instead of 6 per 64 (the shared-memory pipe delivers 128 B/clk/SM whether or not the lanes read the same address).
__device__ __forceinline__ float ldg_stream5(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

template <int U, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) sim5_kernel(const float* __restrict__ rep, const float* __restrict__ scratch, int hw,
                                                                int N, int C, float* __restrict__ out) {
    constexpr int NH = 3;                      // float4 groups per class half
    __shared__ float4 sp[CSS_D * 2 * NH];      // [256][2 halves][NH]
    for (int i = threadIdx.x; i < CSS_D * 2 * NH; i += WARPS * 32) {
        const int d = i / (2 * NH), g = i - d * (2 * NH);
        sp[i] = reinterpret_cast<const float4*>(scratch)[d * (CSS_CMAX / 4) + g];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ch = lane >> 4, pl = lane & 15;
    const float4* myp = sp + ch * NH;
    const int n_wchunks = (N + 63) / 64;
    for (int wc = blockIdx.x * WARPS + warp; wc < n_wchunks; wc += gridDim.x * WARPS) {
        const float* x[4];
        int pix[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            pix[j] = wc * 64 + j * 16 + pl;
            const int p = min(pix[j], N - 1);
            const int b = p / hw;
            x[j] = rep + (size_t)b * CSS_D * hw + (p - b * hw);
        }
        float2 acc[4][2 * NH];
        float n2[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            n2[j] = 0.f;
#pragma unroll
            for (int i = 0; i < 2 * NH; ++i) acc[j][i] = make_float2(0.f, 0.f);
        }
#pragma unroll 1
        for (int d0 = 0; d0 < CSS_D; d0 += U) {
            float v[U][4];
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int j = 0; j < 4; ++j) v[u][j] = ldg_stream5(x[j] + (size_t)(d0 + u) * hw);
#pragma unroll
            for (int u = 0; u < U; ++u) {
#pragma unroll
                for (int j = 0; j < 4; ++j) n2[j] = fmaf(v[u][j], v[u][j], n2[j]);
#pragma unroll
                for (int g = 0; g < NH; ++g) {
                    const float4 q = myp[(d0 + u) * 2 * NH + g];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 vv = make_float2(v[u][j], v[u][j]);
                        acc[j][2 * g + 0] = __ffma2_rn(vv, make_float2(q.x, q.y), acc[j][2 * g + 0]);
                        acc[j][2 * g + 1] = __ffma2_rn(vv, make_float2(q.z, q.w), acc[j][2 * g + 1]);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (pix[j] >= N) continue;
            const int b = pix[j] / hw, s = pix[j] - b * hw;
            const float inv = 1.f / fmaxf(sqrtf(n2[j]), 1e-12f);
            float* o = out + (size_t)b * C * hw + s;
#pragma unroll
            for (int i = 0; i < 2 * NH; ++i) {
                const int c0 = ch * 4 * NH + 2 * i;
                if (c0 < C) o[(size_t)c0 * hw] = acc[j][i].x * inv;
                if (c0 + 1 < C) o[(size_t)(c0 + 1) * hw] = acc[j][i].y * inv;
            }
        }
    }
}

template <int U, int WARPS, int MINB>
static int launch5(const float* rep, const float* scratch, int hw, int N, int C, float* out, cudaStream_t st) {
    const int n_blocks = ((N + 63) / 64 + WARPS - 1) / WARPS;
    const int grid = n_blocks < 148 * MINB ? n_blocks : 148 * MINB;
    sim5_kernel<U, WARPS, MINB><<<grid, WARPS * 32, 0, st>>>(rep, scratch, hw, N, C, out);
    return 0;
}

extern "C" int dev_sim3(int variant, const float* rep, const float* scratch, int hw, int N, int C, float* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    switch (variant) {
        case 0: rc = launch<64, 32, 4, 3, 4>(rep, scratch, hw, N, C, out, st); break;
        case 1: rc = launch<128, 32, 4, 2, 2>(rep, scratch, hw, N, C, out, st); break;
        case 2: rc = launch<128, 32, 3, 2, 4>(rep, scratch, hw, N, C, out, st); break;
        case 3: rc = launch<64, 32, 4, 3, 2>(rep, scratch, hw, N, C, out, st); break;
        case 4: rc = launch<64, 32, 4, 3, 8>(rep, scratch, hw, N, C, out, st); break;
        case 5: rc = launch<64, 16, 6, 3, 4>(rep, scratch, hw, N, C, out, st); break;
        case 6: rc = launch<64, 32, 4, 3, 4, 1>(rep, scratch, hw, N, C, out, st); break;   // loads only
        case 7: rc = launch<64, 32, 4, 3, 4, 2>(rep, scratch, hw, N, C, out, st); break;   // compute only
        case 8: rc = launch4<64, 32, 4, 4, 4>(rep, scratch, hw, N, C, out, st); break;
        case 9: rc = launch4<64, 32, 4, 4, 2>(rep, scratch, hw, N, C, out, st); break;
        case 10: rc = launch4<128, 32, 4, 2, 2>(rep, scratch, hw, N, C, out, st); break;
        case 11: rc = launch4<128, 32, 4, 2, 4>(rep, scratch, hw, N, C, out, st); break;
        case 12: rc = launch4<64, 32, 4, 4, 4, 2>(rep, scratch, hw, N, C, out, st); break;   // compute only
        case 13: rc = launch4<64, 16, 6, 5, 4>(rep, scratch, hw, N, C, out, st); break;
        case 14: rc = launch5<8, 4, 4>(rep, scratch, hw, N, C, out, st); break;
        case 15: rc = launch5<8, 2, 8>(rep, scratch, hw, N, C, out, st); break;
        case 16: rc = launch5<4, 4, 4>(rep, scratch, hw, N, C, out, st); break;
        case 17: rc = launch5<16, 4, 3>(rep, scratch, hw, N, C, out, st); break;
        case 18: rc = launch5<8, 4, 3>(rep, scratch, hw, N, C, out, st); break;
        case 19: rc = launch5<8, 8, 2>(rep, scratch, hw, N, C, out, st); break;
        default: return -1;
    }
    if (rc) return rc;
    return (int)cudaGetLastError();
}
