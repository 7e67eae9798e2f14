// Development microbenchmark: achievable read bandwidth of an NCHW [B,256,h*w] fp32 map when every pixel needs all 256
// channels (thread = pixel, loop over planes), for different CTA widths / unroll depths / per-thread pixel counts.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
__device__ __forceinline__ float ldg_stream(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
template <int THREADS, int U, int PPT>
__global__ void __launch_bounds__(THREADS) k(const float* __restrict__ rep, int hw, int N, float* __restrict__ out) {
    const int n_chunks = (N + THREADS * PPT - 1) / (THREADS * PPT);
    for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const float* x[PPT];
        float acc[PPT];
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            int p = chunk * THREADS * PPT + j * THREADS + threadIdx.x;
            p = min(p, N - 1);
            const int b = p / hw;
            x[j] = rep + (size_t)b * 256 * hw + (p - b * hw);
            acc[j] = 0.f;
        }
#pragma unroll 1
        for (int d0 = 0; d0 < 256; d0 += U) {
            float v[U][PPT];
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int j = 0; j < PPT; ++j) v[u][j] = ldg_stream(x[j] + (size_t)(d0 + u) * hw);
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int j = 0; j < PPT; ++j) acc[j] += v[u][j];
        }
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            const int p = chunk * THREADS * PPT + j * THREADS + threadIdx.x;
            if (p < N) out[p] = acc[j];
        }
    }
}
// channel-split variant: a warp's lanes = 32 pixels, warps of the CTA split the 256 planes (partial sums via smem)
template <int WARPS, int U>
__global__ void __launch_bounds__(WARPS * 32) ksplit(const float* __restrict__ rep, int hw, int N, float* __restrict__ out) {
    __shared__ float part[WARPS][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_chunks = (N + 31) / 32;
    constexpr int DS = 256 / WARPS;
    for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        int p = min(chunk * 32 + lane, N - 1);
        const int b = p / hw;
        const float* x = rep + ((size_t)b * 256 + warp * DS) * hw + (p - b * hw);
        float acc = 0.f;
#pragma unroll 1
        for (int d0 = 0; d0 < DS; d0 += U) {
            float v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = ldg_stream(x + (size_t)(d0 + u) * hw);
#pragma unroll
            for (int u = 0; u < U; ++u) acc += v[u];
        }
        part[warp][lane] = acc;
        __syncthreads();
        if (warp == 0) {
            float s = 0.f;
            for (int w = 0; w < WARPS; ++w) s += part[w][lane];
            if (chunk * 32 + lane < N) out[chunk * 32 + lane] = s;
        }
        __syncthreads();
    }
}
#define RUN(NAME, GRID, BLOCK, ...)                                                                     \
    for (int r = 0; r < 3; ++r) {                                                                       \
        cudaEventRecord(e0);                                                                            \
        __VA_ARGS__<<<GRID, BLOCK>>>(rep[r % 3], hw, N, out);                                           \
        cudaEventRecord(e1); cudaEventSynchronize(e1);                                                  \
        float ms; cudaEventElapsedTime(&ms, e0, e1);                                                    \
        if (r == 2) printf("%-34s %7.1f us  %6.2f TB/s (%s)\n", NAME, ms * 1e3, bytes / ms / 1e9, cudaGetErrorString(cudaGetLastError())); \
    }
int main(int argc, char** argv) {
    const int B = argc > 1 ? atoi(argv[1]) : 16, hw = 81 * 81, N = B * hw;
    const double bytes = (double)N * 1024;
    float* rep[3]; float* out;
    for (int i = 0; i < 3; ++i) { cudaMalloc(&rep[i], (size_t)N * 1024); cudaMemset(rep[i], 0, (size_t)N * 1024); }
    cudaMalloc(&out, N * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("B=%d N=%d (%.0f MB)\n", B, N, bytes / 1e6);
    RUN("thread=px 128thr U8", (N + 127) / 128, 128, k<128, 8, 1>)
    RUN("thread=px 128thr U32", (N + 127) / 128, 128, k<128, 32, 1>)
    RUN("thread=px 256thr U16", (N + 255) / 256, 256, k<256, 16, 1>)
    RUN("thread=px 512thr U16", (N + 511) / 512, 512, k<512, 16, 1>)
    RUN("thread=px 1024thr U8", (N + 1023) / 1024, 1024, k<1024, 8, 1>)
    RUN("thread=2px 128thr U16", (N + 255) / 256, 128, k<128, 16, 2>)
    RUN("thread=4px 128thr U8", (N + 511) / 512, 128, k<128, 8, 4>)
    RUN("ksplit 8 warps U32", (N + 31) / 32, 256, ksplit<8, 32>)
    RUN("ksplit 8 warps U16", (N + 31) / 32, 256, ksplit<8, 16>)
    RUN("ksplit 4 warps U32", (N + 31) / 32, 128, ksplit<4, 32>)
    RUN("ksplit 16 warps U16", (N + 31) / 32, 512, ksplit<16, 16>)
    RUN("ksplit 8 warps U32 persistent", 148 * 8, 256, ksplit<8, 32>)
    return 0;
}
