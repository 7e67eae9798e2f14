// Development harness: one CTA, D[128 x 32] (TMEM) = sum over k-steps of A[128 x 8] . B[32 x 8]^T with tcgen05.mma kind::tf32, for
// several shared-memory layouts of A (K-major / MN-major, SWIZZLE_128B / none), checked against the host.  Used to pin the
// descriptor conventions of css_sim_tc.cu on real hardware.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o dev_umma dev_umma.cu && ./dev_umma
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define M 128
#define N 32
#define KTOT 64          // 8 k-steps of 8

struct Variant {
    const char* name;
    int a_mn_major;          // instruction descriptor bit 15
    int a_layout;            // 0 K-major SW128, 1 MN-major SW128, 2 MN-major no swizzle, 3 K-major no swizzle
    uint32_t a_lbo, a_sbo, a_swz;      // descriptor fields (bytes, layout type)
    uint32_t a_kstep;        // bytes added to the start address per k-step (within a 32-channel block for K-major SW128)
};

__host__ __device__ inline uint32_t a_offset(int layout, int mn, int k) {
    switch (layout) {
        case 0:  return (k / 32) * (M / 8 * 1024) + (mn / 8) * 1024 + (mn % 8) * 128 + ((((k % 32) / 4) ^ (mn % 8)) * 16) + (k % 4) * 4;
        case 1:  return (k / 8) * (M / 32 * 1024) + (mn / 32) * 1024 + (k % 8) * 128 + ((((mn % 32) / 4) ^ (k % 8)) * 16) + (mn % 4) * 4;
        case 2:  return (k / 8) * (M / 4 * 128) + (mn / 4) * 128 + (k % 8) * 16 + (mn % 4) * 4;
        default: return (k / 8) * (M / 8 * 256) + (mn / 8) * 256 + ((k % 8) / 4) * 128 + (mn % 8) * 16 + (k % 4) * 4;
    }
}
// B always K-major SW128: 32 rows x K
__host__ __device__ inline uint32_t b_offset(int n, int k) {
    return (k / 32) * (N / 8 * 1024) + (n / 8) * 1024 + (n % 8) * 128 + ((((k % 32) / 4) ^ (n % 8)) * 16) + (k % 4) * 4;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t swz) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)swz << 61;
    return d;
}

__global__ void __launch_bounds__(128) umma_test(const char* a_img, const char* b_img, int a_bytes, int b_bytes, uint32_t idesc, uint32_t a_lbo,
                                                 uint32_t a_sbo, uint32_t a_swz, uint32_t a_kstep, int a_layout, float* out, uint32_t* dbg) {
    extern __shared__ char raw[];
    char* smem = (char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    char* sA = smem;
    char* sB = smem + 65536;
    __shared__ unsigned long long bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < a_bytes / 4; i += 128) ((uint32_t*)sA)[i] = ((const uint32_t*)a_img)[i];
    for (int i = tid; i < b_bytes / 4; i += 128) ((uint32_t*)sB)[i] = ((const uint32_t*)b_img)[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    if (tid == 0) {
        dbg[0] = tmem;
        for (int ks = 0; ks < KTOT / 8; ++ks) {
            uint32_t a_addr;
            if (a_layout == 0) a_addr = smem_u32(sA) + (ks / 4) * (M / 8 * 1024) + (ks % 4) * 32;
            else a_addr = smem_u32(sA) + ks * a_kstep;
            const uint32_t b_addr = smem_u32(sB) + (ks / 4) * (N / 8 * 1024) + (ks % 4) * 32;
            const uint64_t da = make_desc(a_addr, a_lbo, a_sbo, a_swz), db = make_desc(b_addr, 16, 1024, 2);
            const uint32_t acc = ks > 0;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    {
        uint32_t done = 0;
        const uint32_t b32 = smem_u32(&bar);
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(b32) : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) out[tid * 32 + j] = __uint_as_float(r[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem));
}

int main() {
    static float A[M][KTOT], B[N][KTOT], ref[M][N];
    srand(7);
    for (int m = 0; m < M; ++m)
        for (int k = 0; k < KTOT; ++k) A[m][k] = (float)((rand() % 17) - 8);          // small integers: exact in tf32
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < KTOT; ++k) B[n][k] = (float)((rand() % 9) - 4);
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            float s = 0;
            for (int k = 0; k < KTOT; ++k) s += A[m][k] * B[n][k];
            ref[m][n] = s;
        }
    Variant vs[] = {
        {"A K-major  SW128 (lbo 16, sbo 1024)", 0, 0, 16, 1024, 2, 0},
        {"A MN-major SW128 (lbo 1024, sbo 4096)", 1, 1, 1024, 4096, 2, 4096},
        {"A MN-major SW128 (lbo 4096, sbo 1024) [swapped]", 1, 1, 4096, 1024, 2, 4096},
        {"A MN-major none  (sbo 128 = mn, lbo 4096 = k)", 1, 2, 4096, 128, 0, 4096},
        {"A MN-major none  (lbo 128 = mn, sbo 4096 = k) [swapped]", 1, 2, 128, 4096, 0, 4096},
        {"A K-major  none  (lbo 128, sbo 256)", 0, 3, 128, 256, 0, 4096},
    };
    char *a_img, *b_img, *da, *db;
    float *dout, hout[M * N];
    uint32_t* ddbg;
    a_img = (char*)calloc(65536, 1);
    b_img = (char*)calloc(65536, 1);
    cudaMalloc(&da, 65536);
    cudaMalloc(&db, 65536);
    cudaMalloc(&dout, sizeof(hout));
    cudaMalloc(&ddbg, 64);
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < KTOT; ++k) memcpy(b_img + b_offset(n, k), &B[n][k], 4);
    cudaMemcpy(db, b_img, 65536, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(umma_test, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024);
    {   // does the tensor core TRUNCATE or ROUND an fp32 bit pattern to TF32?  A[0][0] = 1 + 2^-11 + 2^-12, B[0][0] = 1, rest 0
        static float A2[M][KTOT], B2[N][KTOT];
        memset(A2, 0, sizeof(A2));
        memset(B2, 0, sizeof(B2));
        A2[0][0] = 1.0f + 1.0f / 2048 + 1.0f / 4096;
        B2[0][0] = 1.0f;
        A2[1][0] = 1.0f;
        B2[1][0] = 1.0f + 1.0f / 2048 + 1.0f / 4096;
        memset(a_img, 0, 65536);
        memset(b_img, 0, 65536);
        for (int m = 0; m < M; ++m)
            for (int k = 0; k < KTOT; ++k) memcpy(a_img + a_offset(0, m, k), &A2[m][k], 4);
        for (int n = 0; n < N; ++n)
            for (int k = 0; k < KTOT; ++k) memcpy(b_img + b_offset(n, k), &B2[n][k], 4);
        cudaMemcpy(da, a_img, 65536, cudaMemcpyHostToDevice);
        cudaMemcpy(db, b_img, 65536, cudaMemcpyHostToDevice);
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
        umma_test<<<1, 128, 140 * 1024>>>(da, db, 65536, 65536, idesc, 16, 1024, 2, 0, 0, dout, ddbg);
        cudaDeviceSynchronize();
        cudaMemcpy(hout, dout, sizeof(hout), cudaMemcpyDeviceToHost);
        printf("TF32 operand conversion: A=1+2^-11+2^-12 times 1 -> %.10f (1.0 = truncated, 1.0009765625 = rounded); 1 times B=1+2^-11+2^-12 -> %.10f\n",
               hout[0 * N + 0], hout[1 * N + 1]);
        for (int n = 0; n < N; ++n)
            for (int k = 0; k < KTOT; ++k) memcpy(b_img + b_offset(n, k), &B[n][k], 4);
        cudaMemcpy(db, b_img, 65536, cudaMemcpyHostToDevice);
    }
    for (auto& v : vs) {
        memset(a_img, 0, 65536);
        for (int m = 0; m < M; ++m)
            for (int k = 0; k < KTOT; ++k) memcpy(a_img + a_offset(v.a_layout, m, k), &A[m][k], 4);
        cudaMemcpy(da, a_img, 65536, cudaMemcpyHostToDevice);
        cudaMemset(dout, 0xff, sizeof(hout));
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)v.a_mn_major << 15) | (0u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
        umma_test<<<1, 128, 140 * 1024>>>(da, db, 65536, 65536, idesc, v.a_lbo, v.a_sbo, v.a_swz, v.a_kstep, v.a_layout, dout, ddbg);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(hout, dout, sizeof(hout), cudaMemcpyDeviceToHost);
        double maxerr = 0;
        int bad = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                const double d = fabs((double)hout[m * N + n] - ref[m][n]);
                if (!(d <= 1e-3)) ++bad;
                if (d > maxerr || d != d) maxerr = d;
            }
        printf("%-58s: %s  max|err| %.3g  wrong %d / %d   out[0][0..3] = %g %g %g %g  (ref %g %g %g %g)\n", v.name, cudaGetErrorString(e), maxerr,
               bad, M * N, hout[0], hout[1], hout[2], hout[3], ref[0][0], ref[0][1], ref[0][2], ref[0][3]);
        if (e != cudaSuccess) break;
    }
    return 0;
}
