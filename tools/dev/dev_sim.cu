// Development harness: variants of the K1 similarity kernel, timed by tools/dev/run_sim.py.  Not part of the product library.
#include <cuda_runtime.h>
#include <stdint.h>
#define CSS_D 256
#define CSS_CMAX 32

__device__ __forceinline__ float ldg_stream(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

template <int NG, int PPT, int THREADS, int U, bool F2, bool PRE>
__global__ void __launch_bounds__(THREADS) sim_kernel(const float* __restrict__ rep, const float* __restrict__ scratch, int hw, int N,
                                                      int C, float* __restrict__ out) {
    __shared__ float4 sp[CSS_D * NG];
    for (int i = threadIdx.x; i < CSS_D * NG; i += THREADS) {
        int d = i / NG, g = i - d * NG;
        sp[i] = reinterpret_cast<const float4*>(scratch)[d * (CSS_CMAX / 4) + g];
    }
    __syncthreads();
    constexpr int CHUNK = THREADS * PPT;
    const int n_chunks = (N + CHUNK - 1) / CHUNK;
    for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const float* x[PPT];
        int pix[PPT];
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            pix[j] = chunk * CHUNK + j * THREADS + threadIdx.x;
            const int p = min(pix[j], N - 1);
            const int b = p / hw;
            x[j] = rep + (size_t)b * CSS_D * hw + (p - b * hw);
        }
        float2 acc[PPT][2 * NG];
        float n2[PPT];
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            n2[j] = 0.f;
#pragma unroll
            for (int i = 0; i < 2 * NG; ++i) acc[j][i] = make_float2(0.f, 0.f);
        }
        float v[U][PPT], vn[U][PPT];
        if (PRE) {
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int j = 0; j < PPT; ++j) vn[u][j] = ldg_stream(x[j] + (size_t)u * hw);
        }
#pragma unroll 1
        for (int d0 = 0; d0 < CSS_D; d0 += U) {
            if (PRE) {
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int j = 0; j < PPT; ++j) v[u][j] = vn[u][j];
                if (d0 + U < CSS_D) {
#pragma unroll
                    for (int u = 0; u < U; ++u)
#pragma unroll
                        for (int j = 0; j < PPT; ++j) vn[u][j] = ldg_stream(x[j] + (size_t)(d0 + U + u) * hw);
                }
            } else {
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int j = 0; j < PPT; ++j) v[u][j] = ldg_stream(x[j] + (size_t)(d0 + u) * hw);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
#pragma unroll
                for (int j = 0; j < PPT; ++j) n2[j] = fmaf(v[u][j], v[u][j], n2[j]);
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const float4 q = sp[(d0 + u) * NG + g];
#pragma unroll
                    for (int j = 0; j < PPT; ++j) {
                        if (F2) {
                            const float2 vv = make_float2(v[u][j], v[u][j]);
                            acc[j][2 * g + 0] = __ffma2_rn(vv, make_float2(q.x, q.y), acc[j][2 * g + 0]);
                            acc[j][2 * g + 1] = __ffma2_rn(vv, make_float2(q.z, q.w), acc[j][2 * g + 1]);
                        } else {
                            acc[j][2 * g + 0].x = fmaf(v[u][j], q.x, acc[j][2 * g + 0].x);
                            acc[j][2 * g + 0].y = fmaf(v[u][j], q.y, acc[j][2 * g + 0].y);
                            acc[j][2 * g + 1].x = fmaf(v[u][j], q.z, acc[j][2 * g + 1].x);
                            acc[j][2 * g + 1].y = fmaf(v[u][j], q.w, acc[j][2 * g + 1].y);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            if (pix[j] >= N) continue;
            const int b = pix[j] / hw, s = pix[j] - b * hw;
            const float nrm = fmaxf(sqrtf(n2[j]), 1e-12f);
            float* o = out + (size_t)b * C * hw + s;
#pragma unroll
            for (int c = 0; c < 2 * NG; ++c) {
                if (2 * c < C) o[(size_t)(2 * c) * hw] = __fdiv_rn(acc[j][c].x, nrm);
                if (2 * c + 1 < C) o[(size_t)(2 * c + 1) * hw] = __fdiv_rn(acc[j][c].y, nrm);
            }
        }
    }
}

template <int PPT, int THREADS, int U, bool F2, bool PRE>
static void launch(const float* rep, const float* scratch, int hw, int N, int C, float* out, int bps, cudaStream_t st) {
    constexpr int CHUNK = THREADS * PPT;
    int n_chunks = (N + CHUNK - 1) / CHUNK;
    int grid = n_chunks < 148 * bps ? n_chunks : 148 * bps;
    sim_kernel<6, PPT, THREADS, U, F2, PRE><<<grid, THREADS, 0, st>>>(rep, scratch, hw, N, C, out);
}

extern "C" int dev_sim(int variant, const float* rep, const float* scratch, int hw, int N, int C, float* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    switch (variant) {
        case 0: launch<1, 128, 8, false, false>(rep, scratch, hw, N, C, out, 1000, st); break;   // ~ original
        case 1: launch<4, 64, 8, false, false>(rep, scratch, hw, N, C, out, 8, st); break;       // current
        case 2: launch<2, 128, 8, true, false>(rep, scratch, hw, N, C, out, 8, st); break;
        case 3: launch<4, 64, 4, true, false>(rep, scratch, hw, N, C, out, 8, st); break;
        case 4: launch<4, 128, 4, true, false>(rep, scratch, hw, N, C, out, 4, st); break;
        case 5: launch<2, 64, 8, true, false>(rep, scratch, hw, N, C, out, 16, st); break;
        case 6: launch<1, 128, 8, true, false>(rep, scratch, hw, N, C, out, 1000, st); break;
        case 7: launch<2, 128, 4, true, true>(rep, scratch, hw, N, C, out, 8, st); break;
        case 8: launch<4, 64, 4, true, true>(rep, scratch, hw, N, C, out, 8, st); break;
        case 9: launch<2, 64, 4, true, true>(rep, scratch, hw, N, C, out, 16, st); break;
        case 10: launch<1, 128, 8, true, true>(rep, scratch, hw, N, C, out, 1000, st); break;
        case 11: launch<2, 256, 4, true, false>(rep, scratch, hw, N, C, out, 4, st); break;
        default: return -1;
    }
    return (int)cudaGetLastError();
}
