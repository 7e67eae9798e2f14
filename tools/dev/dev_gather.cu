// Development microbenchmark: achievable random 1 KB row-gather bandwidth (L2-resident table) on B200.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

template <int ROWS_IN_FLIGHT>   // per 8-lane group
__global__ void __launch_bounds__(128) gather_kernel(const float4* __restrict__ rows, const int* __restrict__ idx, int n_per_warp,
                                                     float* __restrict__ out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int grp = lane >> 3, l8 = lane & 7;
    const int* my = idx + (size_t)warp * n_per_warp;
    float acc = 0.f;
    for (int base = 0; base < n_per_warp; base += 4 * ROWS_IN_FLIGHT) {
        float4 r[ROWS_IN_FLIGHT][8];
#pragma unroll
        for (int t = 0; t < ROWS_IN_FLIGHT; ++t) {
            const int row = my[base + t * 4 + grp];
            const float4* p = rows + (size_t)row * 64 + l8;
#pragma unroll
            for (int i = 0; i < 8; ++i) r[t][i] = __ldg(p + i * 8);
        }
#pragma unroll
        for (int t = 0; t < ROWS_IN_FLIGHT; ++t)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc += r[t][i].x + r[t][i].y + r[t][i].z + r[t][i].w;
    }
    if (acc == 123.456f) out[0] = acc;
}

int main(int argc, char** argv) {
    const int n_rows = argc > 1 ? atoi(argv[1]) : 104976;      // 107 MB table
    const int n_warps = 5376, n_per_warp = 512;
    float4* rows; int* idx; float* out;
    cudaMalloc(&rows, (size_t)n_rows * 1024);
    cudaMemset(rows, 0, (size_t)n_rows * 1024);
    cudaMalloc(&idx, (size_t)n_warps * n_per_warp * 4);
    cudaMalloc(&out, 4);
    int* h = (int*)malloc((size_t)n_warps * n_per_warp * 4);
    srand(1);
    for (size_t i = 0; i < (size_t)n_warps * n_per_warp; ++i) h[i] = (int)(((uint64_t)rand() * 2654435761u) % n_rows);
    cudaMemcpy(idx, h, (size_t)n_warps * n_per_warp * 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double bytes = (double)n_warps * n_per_warp * 1024;
    for (int variant = 0; variant < 3; ++variant) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (variant == 0) gather_kernel<1><<<n_warps / 4, 128>>>(rows, idx, n_per_warp, out);
            if (variant == 1) gather_kernel<2><<<n_warps / 4, 128>>>(rows, idx, n_per_warp, out);
            if (variant == 2) gather_kernel<4><<<n_warps / 4, 128>>>(rows, idx, n_per_warp, out);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep == 2) printf("rows=%d (%.0f MB) rows_in_flight/group=%d: %.1f us  %.2f TB/s  (%s)\n", n_rows, n_rows / 1024.0, 1 << variant,
                                 ms * 1e3, bytes / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
