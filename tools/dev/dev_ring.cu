// Development microbenchmark: can ONE 512-thread CTA per SM (the producer shape of css_sim_tc.cu: lane = pixel, warp = (pixel
// group, channel octet), 8 loads per 32-channel chunk) keep an NCHW map streaming from HBM with a register ring of R chunks?
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
__device__ __forceinline__ float ldg_stream(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
template <int R, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) ring_kernel(const float* __restrict__ rep, int hw, int N, float* __restrict__ out) {
    constexpr int WARPS = THREADS / 32, OCT = WARPS / 4, U = 32 / OCT;      // channels per thread per 32-channel chunk
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, pg = warp & 3, co = warp >> 2, m = pg * 32 + lane;
    const int n_tiles = (N + 127) / 128, n_chunks_total = 8;
    float ring[R][U];
    float acc = 0.f;
    auto ptr = [&](int tile, int ch) {
        const int p = min(tile * 128 + m, N - 1), b = p / hw;
        return rep + ((size_t)b * 256 + ch * 32 + co * U) * hw + (p - b * hw);
    };
    // flat chunk index g over this CTA's tiles
    const int my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x, total = my_tiles * n_chunks_total;
#pragma unroll
    for (int i = 0; i < R; ++i) {
        if (i < total) {
            const float* p = ptr(blockIdx.x + (i / 8) * gridDim.x, i % 8);
#pragma unroll
            for (int u = 0; u < U; ++u) ring[i][u] = ldg_stream(p + (size_t)u * hw);
        }
    }
    for (int g0 = 0; g0 < total; g0 += R) {
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int g = g0 + i;
            if (g < total) {
#pragma unroll
                for (int u = 0; u < U; ++u) acc += ring[i][u];
                const int gn = g + R;
                if (gn < total) {
                    const float* p = ptr(blockIdx.x + (gn / 8) * gridDim.x, gn % 8);
#pragma unroll
                    for (int u = 0; u < U; ++u) ring[i][u] = ldg_stream(p + (size_t)u * hw);
                }
            }
        }
    }
    if (acc == 123.456f) out[0] = acc;
}
#define RUN(NAME, ...)                                                                                   \
    for (int r = 0; r < 3; ++r) {                                                                        \
        cudaEventRecord(e0);                                                                             \
        __VA_ARGS__;                                                                                     \
        cudaEventRecord(e1); cudaEventSynchronize(e1);                                                   \
        float ms; cudaEventElapsedTime(&ms, e0, e1);                                                     \
        if (r == 2) printf("%-44s %7.1f us  %6.2f TB/s (%s)\n", NAME, ms * 1e3, bytes / ms / 1e9, cudaGetErrorString(cudaGetLastError())); \
    }
int main(int argc, char** argv) {
    const int B = argc > 1 ? atoi(argv[1]) : 16, hw = 81 * 81, N = B * hw;
    const double bytes = (double)N * 1024;
    float* rep[3]; float* out;
    for (int i = 0; i < 3; ++i) { cudaMalloc(&rep[i], (size_t)N * 1024); cudaMemset(rep[i], 0, (size_t)N * 1024); }
    cudaMalloc(&out, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("B=%d N=%d (%.0f MB), one CTA per SM\n", B, N, bytes / 1e6);
    RUN("512 thr, ring 2 x 8 loads (16 KB... in flight)", ring_kernel<2, 512><<<148, 512>>>(rep[r % 3], hw, N, out))
    RUN("512 thr, ring 4 x 8 loads (64 KB in flight)", ring_kernel<4, 512><<<148, 512>>>(rep[r % 3], hw, N, out))
    RUN("512 thr, ring 8 x 8 loads (128 KB in flight)", ring_kernel<8, 512><<<148, 512>>>(rep[r % 3], hw, N, out))
    RUN("512 thr, ring 12 x 8 loads (192 KB in flight)", ring_kernel<12, 512><<<148, 512>>>(rep[r % 3], hw, N, out))
    RUN("1024 thr, ring 4 x 4 loads (64 KB in flight)", ring_kernel<4, 1024><<<148, 1024>>>(rep[r % 3], hw, N, out))
    RUN("1024 thr, ring 8 x 4 loads (128 KB in flight)", ring_kernel<8, 1024><<<148, 1024>>>(rep[r % 3], hw, N, out))
    RUN("1024 thr, ring 12 x 4 loads (192 KB in flight)", ring_kernel<12, 1024><<<148, 1024>>>(rep[r % 3], hw, N, out))
    return 0;
}
