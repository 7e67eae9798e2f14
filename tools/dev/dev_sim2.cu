// Development harness 2: K-split-across-lanes variants of the K1 similarity kernel.  Not part of the product library.
#include <cuda_runtime.h>
#include <stdint.h>
#define CSS_D 256
#define CSS_CMAX 32

__device__ __forceinline__ float ldg_stream(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// lane = ks * (32/KS) + psub ; lane handles PPT pixels x (256/KS) channels
template <int NG, int PPT, int KS, int U, int WARPS, int MINB, bool IL>
__global__ void __launch_bounds__(WARPS * 32, MINB) sim_kernel(const float* __restrict__ rep, const float* __restrict__ scratch, int hw,
                                                               int N, int C, float* __restrict__ out) {
    constexpr int DS = CSS_D / KS;            // channels per slice
    constexpr int PL = 32 / KS;               // pixel lanes per warp
    constexpr int WP = PL * PPT;              // pixels per warp
    constexpr int SL = DS * NG + 1;           // float4 per slice (+1 skew: slices land in different banks)
    __shared__ float4 sp[KS * SL];
    for (int i = threadIdx.x; i < CSS_D * NG; i += WARPS * 32) {
        int d = i / NG, g = i - d * NG;
        sp[(d / DS) * SL + (d % DS) * NG + g] = reinterpret_cast<const float4*>(scratch)[d * (CSS_CMAX / 4) + g];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ks = lane / PL, psub = lane % PL;
    const float4* myp = sp + ks * SL;
    const int n_wchunks = IL ? (N + WP * WARPS - 1) / (WP * WARPS) : (N + WP - 1) / WP;
    for (int wc = IL ? blockIdx.x : blockIdx.x * WARPS + warp; wc < n_wchunks; wc += IL ? gridDim.x : gridDim.x * WARPS) {
        const float* x[PPT];
        int pix[PPT];
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            pix[j] = IL ? wc * (WP * WARPS) + j * (PL * WARPS) + warp * PL + psub : wc * WP + j * PL + psub;
            const int p = min(pix[j], N - 1);
            const int b = p / hw;
            x[j] = rep + ((size_t)b * CSS_D + ks * DS) * hw + (p - b * hw);
        }
        float2 acc[PPT][2 * NG];
        float n2[PPT];
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            n2[j] = 0.f;
#pragma unroll
            for (int i = 0; i < 2 * NG; ++i) acc[j][i] = make_float2(0.f, 0.f);
        }
#pragma unroll 1
        for (int d0 = 0; d0 < DS; d0 += U) {
            float v[U][PPT];
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int j = 0; j < PPT; ++j) v[u][j] = ldg_stream(x[j] + (size_t)(d0 + u) * hw);
#pragma unroll
            for (int u = 0; u < U; ++u) {
#pragma unroll
                for (int j = 0; j < PPT; ++j) n2[j] = fmaf(v[u][j], v[u][j], n2[j]);
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const float4 q = myp[(d0 + u) * NG + g];
#pragma unroll
                    for (int j = 0; j < PPT; ++j) {
                        const float2 vv = make_float2(v[u][j], v[u][j]);
                        acc[j][2 * g + 0] = __ffma2_rn(vv, make_float2(q.x, q.y), acc[j][2 * g + 0]);
                        acc[j][2 * g + 1] = __ffma2_rn(vv, make_float2(q.z, q.w), acc[j][2 * g + 1]);
                    }
                }
            }
        }
        // combine the KS channel slices (fixed shuffle order -> deterministic)
#pragma unroll
        for (int o = PL; o < 32; o <<= 1) {
#pragma unroll
            for (int j = 0; j < PPT; ++j) {
                n2[j] += __shfl_xor_sync(0xffffffffu, n2[j], o);
#pragma unroll
                for (int i = 0; i < 2 * NG; ++i) {
                    acc[j][i].x += __shfl_xor_sync(0xffffffffu, acc[j][i].x, o);
                    acc[j][i].y += __shfl_xor_sync(0xffffffffu, acc[j][i].y, o);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            if (pix[j] >= N || (j % KS) != ks) continue;      // slice j%KS writes pixel j
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

template <int PPT, int KS, int U, int WARPS, int MINB, bool IL = false>
static void launch(const float* rep, const float* scratch, int hw, int N, int C, float* out, cudaStream_t st) {
    constexpr int WP = (32 / KS) * PPT;
    int n_blocks = ((N + WP - 1) / WP + WARPS - 1) / WARPS;
    int grid = n_blocks < 148 * MINB ? n_blocks : 148 * MINB;
    sim_kernel<6, PPT, KS, U, WARPS, MINB, IL><<<grid, WARPS * 32, 0, st>>>(rep, scratch, hw, N, C, out);
}

extern "C" int dev_sim(int variant, const float* rep, const float* scratch, int hw, int N, int C, float* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    switch (variant) {
        case 0: launch<4, 4, 8, 4, 3>(rep, scratch, hw, N, C, out, st); break;
        case 1: launch<4, 4, 16, 4, 3>(rep, scratch, hw, N, C, out, st); break;
        case 2: launch<2, 4, 8, 4, 5>(rep, scratch, hw, N, C, out, st); break;
        case 3: launch<2, 4, 16, 4, 4>(rep, scratch, hw, N, C, out, st); break;
        case 4: launch<4, 2, 8, 4, 3>(rep, scratch, hw, N, C, out, st); break;
        case 5: launch<2, 2, 16, 4, 4>(rep, scratch, hw, N, C, out, st); break;
        case 6: launch<2, 2, 16, 8, 2, true>(rep, scratch, hw, N, C, out, st); break;
        case 7: launch<2, 4, 16, 8, 2, true>(rep, scratch, hw, N, C, out, st); break;
        case 8: launch<2, 2, 16, 4, 4, true>(rep, scratch, hw, N, C, out, st); break;
        case 9: launch<2, 1, 16, 8, 2, true>(rep, scratch, hw, N, C, out, st); break;
        case 10: launch<2, 4, 32, 4, 3>(rep, scratch, hw, N, C, out, st); break;
        case 11: launch<1, 2, 16, 8, 3, true>(rep, scratch, hw, N, C, out, st); break;
        default: return -1;
    }
    return (int)cudaGetLastError();
}
