// Development microbenchmark: issue rate of scalar FFMA vs packed FFMA2 (fp32x2) on sm_100a, per SM sub-partition.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dev_fma dev_fma.cu && ./dev_fma
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void fma_kernel(float* out, long long* cycles, float seed, int iters) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(seed + i, seed - i);
    const float2 b = make_float2(seed * 0.5f, seed * 0.25f), c = make_float2(seed * 0.125f, seed * 0.0625f);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) {            // 2 scalar FFMA per slot (3 distinct registers each)
                a[i].x = fmaf(a[i].x, b.x, c.x);
                a[i].y = fmaf(a[i].y, b.y, c.y);
            } else {                    // 1 packed FFMA2 per slot
                a[i] = __ffma2_rn(a[i], b, c);
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
    float* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(float));
    cudaMalloc(&cyc, 148 * sizeof(long long));
    const int iters = 4096;
    for (int warps = 4; warps <= 32; warps *= 2) {
        for (int mode = 0; mode < 2; ++mode) {
            if (mode == 0) fma_kernel<0><<<148, warps * 32>>>(out, cyc, 1.0001f, iters);
            else fma_kernel<1><<<148, warps * 32>>>(out, cyc, 1.0001f, iters);
            cudaDeviceSynchronize();
            long long h[148];
            cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double avg = 0;
            for (int i = 0; i < 148; ++i) avg += (double)h[i];
            avg /= 148;
            const double fma_per_thread = (double)iters * 16;            // 16 fp32 FMAs per iteration per thread in both modes
            const double per_clk_sm = fma_per_thread * warps * 32 / avg;
            printf("%s  warps/SM %2d (%d per sub-partition): %.0f cycles, %.1f fp32 FMA/clk/SM\n", mode ? "FFMA2" : "FFMA ", warps, warps / 4, avg,
                   per_clk_sm);
        }
    }
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
