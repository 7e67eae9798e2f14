// Measurement probe (NOT part of the product library): the ceiling of the two access patterns the path is bound by, measured
// live by bench.py next to the kernels it reports, on the same device, table size and clocks.
//   css_probe_gather_gbs : random 1 KB-row gathers from a [n_rows][256] fp32 table, no math, 64 warps per SM -- the L2 -> SM
//                          path the scorer (css_score_ce) is bound by.  bench.py passes the step's own pixel-major copy.
//   css_probe_copy_gbs   : device-to-device copy (read + write bytes) -- cross-check of MEASURED_PEAKS.json's hbm_gbs.
// Built by __graft_entry__.build() into tools/dev/libcss_probe.so; bench.py reports null when it is absent.
#include <cuda_runtime.h>
#include <stdint.h>

template <int RIF>   // rows in flight per 8-lane group
__global__ void __launch_bounds__(128) probe_gather_kernel(const float4* __restrict__ rows, unsigned n_rows, int n_per_warp,
                                                           float* __restrict__ out) {
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const unsigned grp = lane >> 3, l8 = lane & 7;
    float acc = 0.f;
    for (int base = 0; base < n_per_warp; base += 4 * RIF) {
        float4 r[RIF][8];
#pragma unroll
        for (int t = 0; t < RIF; ++t) {
            unsigned h = (warp * 0x9E3779B1u) ^ ((unsigned)(base + t * 4 + grp) * 0x85EBCA77u);     // cheap mix: a different row per group
            h ^= h >> 15; h *= 0xC2B2AE3Du; h ^= h >> 13;
            const float4* p = rows + (size_t)(h % n_rows) * 64 + l8;
#pragma unroll
            for (int i = 0; i < 8; ++i) r[t][i] = __ldg(p + i * 8);
        }
#pragma unroll
        for (int t = 0; t < RIF; ++t)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc += r[t][i].x + r[t][i].y + r[t][i].z + r[t][i].w;
    }
    if (acc == 123.456f) out[0] = acc;
}

extern "C" double css_probe_gather_gbs(const void* table, long long n_rows, int n_warps, int n_per_warp, int reps, void* scratch4) {
    if (!table || n_rows <= 0 || n_warps <= 0 || n_per_warp <= 0 || !scratch4) return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const double bytes = (double)n_warps * n_per_warp * 1024.0;
    double best = 0.0;
    for (int variant = 0; variant < 3; ++variant) {
        for (int r = 0; r < reps + 1; ++r) {
            cudaEventRecord(e0);
            const int grid = (n_warps + 3) / 4;
            if (variant == 0) probe_gather_kernel<1><<<grid, 128>>>((const float4*)table, (unsigned)n_rows, n_per_warp, (float*)scratch4);
            if (variant == 1) probe_gather_kernel<2><<<grid, 128>>>((const float4*)table, (unsigned)n_rows, n_per_warp, (float*)scratch4);
            if (variant == 2) probe_gather_kernel<4><<<grid, 128>>>((const float4*)table, (unsigned)n_rows, n_per_warp, (float*)scratch4);
            cudaEventRecord(e1);
            if (cudaEventSynchronize(e1) != cudaSuccess) return -2.0;
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            if (r > 0 && ms > 0.f && bytes / ms / 1e6 > best) best = bytes / ms / 1e6;     // GB/s; the first run warms L2
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return best;
}

extern "C" double css_probe_copy_gbs(void* dst, const void* src, long long bytes, int reps) {
    if (!dst || !src || bytes <= 0) return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int r = 0; r < reps + 1; ++r) {
        cudaEventRecord(e0);
        cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) return -2.0;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms > 0.f && 2.0 * bytes / ms / 1e6 > best) best = 2.0 * bytes / ms / 1e6;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return best;
}
