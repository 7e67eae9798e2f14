// Shared device/host helpers of libcss_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/css_b200.h"

#define CSS_D 256            // feature width (reference: output_dim=256, mix_label.py:75,93)
#define CSS_CMAX 32          // classes fit one 32-bit set per pixel
#define CSS_SEL_TILE 256     // pixels per selection tile (= threads per block of the selection kernels)

void css_set_error(const char* fmt, ...);

#define CSS_CHECK_ARG(cond, code, ...)                  \
    do {                                                \
        if (!(cond)) {                                  \
            css_set_error(__VA_ARGS__);                 \
            return (code);                              \
        }                                               \
    } while (0)

#define CSS_CHECK_LAUNCH(name, n_kernels)                                              \
    do {                                                                               \
        css_count_launches(n_kernels);                                                 \
        cudaError_t e__ = cudaGetLastError();                                          \
        if (e__ != cudaSuccess) {                                                      \
            css_set_error("%s: %s", name, cudaGetErrorString(e__));                    \
            return (int)e__;                                                           \
        }                                                                              \
    } while (0)

static inline int css_check_dims(int C, int D) {
    if (D != CSS_D) { css_set_error("D must be %d (got %d)", CSS_D, D); return CSS_E_DIM; }
    if (C < 1 || C > CSS_CMAX) { css_set_error("C must be in [1,%d] (got %d)", CSS_CMAX, C); return CSS_E_DIM; }
    return 0;
}

int css_cached_sm_count();
void css_count_launches(int n);     // bookkeeping behind css_launch_count()
bool css_pdl_enabled();             // programmatic dependent launch between the path's kernels (opt-in: CSS_B200_PDL=1 / css_set_pdl)

#ifdef __CUDACC__
// Every kernel of the library is launched through css_launch: same as <<<grid, block, smem, stream>>>, plus the programmatic
// stream serialization attribute, so that a kernel's CTAs are scheduled while the previous kernel of the stream drains instead
// of after it has retired (the path is ~16 short launches per step: the gaps between them are a measurable share of it).
// Every kernel starts with css_pdl_enter(): wait for the previous grid to complete and flush (nothing is read before that), then
// let the next grid start launching.  Both are no-ops for a kernel launched without the attribute.  Captured into CUDA graphs as
// programmatic edges.
template <typename... KArgs, typename... Args>
static inline cudaError_t css_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = css_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
__device__ __forceinline__ void css_pdl_enter() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
#endif

#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming (read-once) loads: keep them out of L1
__device__ __forceinline__ float ldg_stream(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ldg_stream4(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// Pixel-major rows are stored in the representation map's own dtype (fp32, or bf16 = lossless for a bf16 map and half the
// gather bytes).  row_f4(rows, row, q) returns elements 4q..4q+3 of a row widened to fp32.
__device__ __forceinline__ float4 bf16x4_to_f4(uint2 u) {
    return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                       __uint_as_float(u.y & 0xffff0000u));
}
__device__ __forceinline__ float4 row_f4(const float* rows, size_t row, int q) {
    return __ldg(reinterpret_cast<const float4*>(rows) + row * (CSS_D / 4) + q);
}
__device__ __forceinline__ float4 row_f4(const __nv_bfloat16* rows, size_t row, int q) {
    return bf16x4_to_f4(__ldg(reinterpret_cast<const uint2*>(rows) + row * (CSS_D / 4) + q));
}

// Philox4x32-10 (Salmon et al., SC'11), the counter-based generator the device sampler is built on.
struct Philox {
    static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    __device__ __forceinline__ static uint4 run(uint4 c, uint2 k) {
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
            uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
            c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
            k.x += W0;
            k.y += W1;
        }
        return c;
    }
};
#endif
