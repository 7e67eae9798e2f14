// K1 on the 5th-generation tensor cores: the [pixels x 256] . [256 x C] similarity product of the rep pass as tcgen05.mma
// kind::tf32 with the accumulators in tensor memory.  Reference: generalframeworks/networks/ddp_model.py:104-110 (teacher
// cosine map) and :147-154 / :230-237 (student softmax similarity); the pixel-major copy of loss.py:85.
//
// Why: the FFMA2 form of this pass (css_sim.cu) is bound by its multiply side -- 21 class dots per 4-byte element, fed by
// shared-memory broadcasts, with the 42-48 accumulator registers per thread leaving no room to keep the next loads in flight
// (ncu r02a: 16 % warps active, long-scoreboard stalls, 2.0 TB/s).  Here the multiply side moves to the tensor pipe, the
// accumulators to TMEM, and the threads only stream: load, split, stage, transpose.
//
// Exactness: a TF32 operand keeps 11 significant bits, so every fp32 value x is staged twice, hi = rna_tf32(x) and
// lo = rna_tf32(x - hi) (x - hi is exact in fp32), for the map (A) and for the pre-normalised prototypes (B), and the product is
// A_hi.B_hi + A_lo.B_hi + A_hi.B_lo (the dropped lo.lo term is 2^-22 relative).  The tensor core accumulates in fp32 with
// truncation, which biases long chains; the 256-channel contraction is therefore cut into four 64-channel accumulators (8 MMAs
// each) plus one accumulator for the two correction terms, summed by the epilogue in fp32.
//
// Shape of one CTA (512 threads, one per SM, persistent over 128-pixel tiles):
//   * warp w = (pixel group w & 3, channel quarter w >> 2): lane = one pixel, 16 consecutive channels of the current 64-channel
//     chunk: 16 independent 4-byte loads in flight for the NEXT chunk while the current one is split and staged;
//   * A stage in shared memory: K-major SWIZZLE_128B (row = pixel, 128 B = 32 channels, 8-row atoms of 1 KB), the layout the
//     tensor core reads without transposition: a thread's 16 channels are 64 contiguous bytes of its pixel's row, i.e. four
//     128-bit stores per half, and the XOR swizzle spreads the 8 pixels of a store phase over all banks; two stages
//     (hi | lo, 64 KB each).  (An MN-major A descriptor -- pixels contiguous, as the NCHW planes are -- returned zeros for
//     kind::tf32 in every canonical layout tried on B200, tools/dev/dev_umma.cu; K-major is also the cheaper store pattern.)
//   * B (32 class slots x 256 channels, K-major SWIZZLE_128B, hi | lo = 64 KB) is written once per call in exactly the shared
//     memory image by proto_prep_tc_kernel and copied in at kernel start;
//   * thread 0 issues the 24 MMAs of a chunk after the CTA barrier (M = 128, N = 32, K = 8 each) and commits them to the
//     stage's mbarrier; the stage is reused when that barrier completes;
//   * epilogue (warps 0-3, lane = TMEM lane = pixel): tcgen05.ld of the five accumulators, 1 / max(||x||, 1e-12), optional
//     softmax, coalesced NCHW stores; the pixel-major rows leave registers as whole 32-byte sectors, as in css_sim.cu.
#include "css_common.cuh"

#define TC_THREADS 512
#define TC_M 128                      // pixels per tile  (UMMA M)
#define TC_N 32                       // class slots      (UMMA N)
#define TC_KC 64                      // channels per chunk (= per shared-memory stage)
#define TC_NCHUNK (CSS_D / TC_KC)     // 4 chunks per tile, one main accumulator each
#define TC_U 16                       // channels per thread per chunk
#define TC_STAGE_HALF (TC_KC * TC_M * 4)          // 32 KB: hi (or lo) part of one stage
#define TC_B_HALF (TC_N * CSS_D * 4)              // 32 KB: hi (or lo) image of the prototypes
#define TC_TMEM_COLS 256                          // 5 accumulators x 32 columns, rounded up to a power of two
#define TC_SMEM_BYTES (2 * TC_B_HALF + 2 * 2 * TC_STAGE_HALF + 4096 + 1024)   // B | 2 stages | barriers + norm partials | align

// ---------------------------------------------------------------------------------------------------------------------
// shared-memory images
// ---------------------------------------------------------------------------------------------------------------------
// B, K-major SWIZZLE_128B: atom = 8 class rows x 128 B (32 channels); atoms of a 32-channel block stacked along N (4 x 1 KB),
// blocks along K 4 KB apart.  Byte offset of element (class n, channel k):
__host__ __device__ __forceinline__ int tc_b_offset(int n, int k) {
    const int kk = k >> 5, g = n >> 3, r = n & 7, c = (k & 31) >> 2, e = k & 3;
    return kk * 4096 + g * 1024 + r * 128 + ((c ^ r) << 4) + (e << 2);
}

__device__ __forceinline__ uint32_t rna_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

// F.normalize(prototypes, dim=-1) (eps 1e-12, ddp_model.py:107) split into TF32 hi / lo parts, written in the shared-memory image
__global__ void __launch_bounds__(CSS_D) proto_prep_tc_kernel(const float* __restrict__ protos, char* __restrict__ image, int C) {
    __shared__ float part[CSS_D / 32];
    const int c = blockIdx.x, d = threadIdx.x;
    float v = 0.f;
    if (c < C) {
        v = protos[c * CSS_D + d];
        const float s = warp_sum(v * v);
        if ((d & 31) == 0) part[d >> 5] = s;
    }
    __syncthreads();
    uint32_t hi = 0u, lo = 0u;
    if (c < C) {
        float n2 = 0.f;
#pragma unroll
        for (int i = 0; i < CSS_D / 32; ++i) n2 += part[i];
        const float p = __fdiv_rn(v, fmaxf(sqrtf(n2), 1e-12f));
        hi = rna_tf32(p);
        lo = rna_tf32(__fsub_rn(p, __uint_as_float(hi)));
    }
    const int off = tc_b_offset(c, d);
    *reinterpret_cast<uint32_t*>(image + off) = hi;
    *reinterpret_cast<uint32_t*>(image + TC_B_HALF + off) = lo;
}

// ---------------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// UMMA shared-memory descriptor: start address, leading / stride byte offsets (all >> 4), version 1, SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;           // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;           // layout type: SWIZZLE_128B
    return d;
}

// instruction descriptor: D = F32, A = B = TF32, both K-major, N = 32, M = 128
#define TC_IDESC ((1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (0u << 16) | ((TC_N >> 3) << 17) | ((TC_M >> 4) << 24))

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(TC_IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct TcShared {                     // small state behind the big buffers
    unsigned long long mma_done[2];   // stage s may be overwritten: the MMAs that read it have completed
    unsigned long long acc_full;      // the tile's accumulators are complete
    uint32_t tmem_base;
    uint32_t pad;
    float n2[4][TC_M];                // per channel quarter partial ||x||^2
};

// ---------------------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------------------
// NP: 8-column parts of the accumulators the epilogue reads (3 when C <= 24, else 4)
template <bool ROWS, int NP>
__global__ void __launch_bounds__(TC_THREADS, 1) rep_pass_tc_kernel(const float* __restrict__ rep, const char* __restrict__ b_image, int hw,
                                                                    int N, int C, int mode, float temp, float* __restrict__ out,
                                                                    float* __restrict__ rows, float* __restrict__ norms) {
    extern __shared__ char smem_raw[];
    char* smem = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B atoms: 1 KB aligned
    char* sB = smem;                                        // [hi 32 KB | lo 32 KB]
    char* sA = smem + 2 * TC_B_HALF;                        // stage s at s * 64 KB: [hi 32 KB | lo 32 KB]
    TcShared* sh = reinterpret_cast<TcShared*>(sA + 2 * 2 * TC_STAGE_HALF);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int pg = warp & 3, cq = warp >> 2;
    const int m = pg * 32 + lane;                           // pixel of the tile = TMEM lane = row of A

    // ---- one-time set-up: barriers, tensor memory, the prototype image ----
    if (tid == 0) {
        mbar_init(smem_u32(&sh->mma_done[0]), 1);
        mbar_init(smem_u32(&sh->mma_done[1]), 1);
        mbar_init(smem_u32(&sh->acc_full), 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)), "n"(TC_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    {
        const uint4* src = reinterpret_cast<const uint4*>(b_image);
        uint4* dst = reinterpret_cast<uint4*>(sB);
        for (int i = tid; i < 2 * TC_B_HALF / 16; i += TC_THREADS) dst[i] = __ldg(src + i);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sh->tmem_base;
    const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);

    const int n_tiles = (N + TC_M - 1) / TC_M;
    // byte offset of this thread's pixel row inside a stage half (K-major SWIZZLE_128B: 32-channel blocks of 128 x 128 B = 16 KB,
    // 8-row atoms of 1 KB): its 16 channels cq*16.. are the 16-byte chunks (cq & 1)*4 + j of block cq >> 1, XOR-swizzled by m & 7
    const uint32_t a_thread = (uint32_t)((cq >> 1) * (TC_M * 128) + (m >> 3) * 1024 + (m & 7) * 128);
    const int chunk0 = (cq & 1) * 4, r7 = m & 7;

    float cur[TC_U], nxt[TC_U];
    const float* xp = nullptr;                              // first channel of this thread in the current tile
    auto tile_ptr = [&](int tile) -> const float* {
        const int p = min(tile * TC_M + m, N - 1);          // out-of-range lanes re-read the last pixel, never write
        const int b = p / hw;
        return rep + ((size_t)b * CSS_D + cq * TC_U) * hw + (p - b * hw);
    };
    auto load_chunk = [&](const float* base, int ch, float (&v)[TC_U]) {
        const float* p = base + (size_t)(ch * TC_KC) * hw;
#pragma unroll
        for (int u = 0; u < TC_U; ++u, p += hw) v[u] = ldg_stream(p);
    };

    int tile = blockIdx.x;
    if (tile < n_tiles) {
        xp = tile_ptr(tile);
        load_chunk(xp, 0, cur);
    }
    uint32_t g = 0;                                         // chunks this CTA has staged so far (stage = g & 1)
    uint32_t tile_it = 0;
    for (; tile < n_tiles; tile += gridDim.x, ++tile_it) {
        const int pix = tile * TC_M + m;
        float n2 = 0.f;
        const int next_tile = tile + gridDim.x;
        const float* xp_next = next_tile < n_tiles ? tile_ptr(next_tile) : nullptr;
#pragma unroll
        for (int ch = 0; ch < TC_NCHUNK; ++ch, ++g) {
            // 1. the next chunk's loads go out before anything else touches the current one
            if (ch + 1 < TC_NCHUNK) load_chunk(xp, ch + 1, nxt);
            else if (xp_next) load_chunk(xp_next, 0, nxt);
            // 2. the stage is free once the MMAs of its previous use have completed
            const uint32_t stage = g & 1u;
            if (g >= 2) mbar_wait(smem_u32(&sh->mma_done[stage]), ((g >> 1) - 1u) & 1u);
            // 3. split into TF32 hi / lo and stage K-major (the thread's 16 channels = four 16-byte chunks of its pixel's row)
            char* st_hi = sA + stage * (2 * TC_STAGE_HALF) + a_thread;
            char* st_lo = st_hi + TC_STAGE_HALF;
#pragma unroll
            for (int j = 0; j < TC_U / 4; ++j) {
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float x = cur[4 * j + e];
                    hi[e] = rna_tf32(x);
                    lo[e] = rna_tf32(__fsub_rn(x, __uint_as_float(hi[e])));
                    n2 = fmaf(x, x, n2);
                }
                const uint32_t off = (uint32_t)(((chunk0 + j) ^ r7) << 4);
                *reinterpret_cast<uint4*>(st_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(st_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
            // 4. pixel-major rows: a lane holds 64 contiguous bytes of its pixel's row; lane pairs swap 16-byte chunks so that
            //    every 128-bit store instruction completes whole 32-byte sectors
            if (ROWS) {
                const int odd = lane & 1;
                const int pA = pix - odd, pB = pA + 1;
                float* rA = rows + (size_t)pA * CSS_D + ch * TC_KC + cq * TC_U + 4 * odd;
                float* rB = rA + CSS_D;
#pragma unroll
                for (int hh = 0; hh < TC_U / 8; ++hh) {
                    const int ue = 8 * hh, uo = 8 * hh + 4;
                    const float4 own_e = make_float4(cur[ue], cur[ue + 1], cur[ue + 2], cur[ue + 3]);
                    const float4 own_o = make_float4(cur[uo], cur[uo + 1], cur[uo + 2], cur[uo + 3]);
                    const float4 snd = odd ? own_e : own_o;
                    float4 rcv;
                    rcv.x = __shfl_xor_sync(0xffffffffu, snd.x, 1);
                    rcv.y = __shfl_xor_sync(0xffffffffu, snd.y, 1);
                    rcv.z = __shfl_xor_sync(0xffffffffu, snd.z, 1);
                    rcv.w = __shfl_xor_sync(0xffffffffu, snd.w, 1);
                    const float4 first = odd ? rcv : own_e;
                    const float4 second = odd ? own_o : rcv;
                    if (pA < N) *reinterpret_cast<float4*>(rA + 8 * hh) = first;
                    if (pB < N) *reinterpret_cast<float4*>(rB + 8 * hh) = second;
                }
            }
            if (ch == TC_NCHUNK - 1) sh->n2[cq][m] = n2;
            // 5. make the staged operands visible to the tensor core (async proxy), then one thread issues the chunk's MMAs
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                const uint32_t a_hi = sA_u + stage * (2 * TC_STAGE_HALF), a_lo = a_hi + TC_STAGE_HALF;
#pragma unroll
                for (int ks = 0; ks < TC_KC / 8; ++ks) {
                    const uint32_t k = (uint32_t)(ch * TC_KC + ks * 8);
                    const uint32_t b_off = (k >> 5) * 4096 + ((k & 31) >> 3) * 32;
                    const uint32_t a_off = (uint32_t)((ks >> 2) * (TC_M * 128) + (ks & 3) * 32);
                    const uint64_t da_hi = umma_desc(a_hi + a_off, 16, 1024), da_lo = umma_desc(a_lo + a_off, 16, 1024);
                    const uint64_t db_hi = umma_desc(sB_u + b_off, 16, 1024), db_lo = umma_desc(sB_u + TC_B_HALF + b_off, 16, 1024);
                    umma_tf32(tmem + ch * TC_N, da_hi, db_hi, ks > 0);                         // main: 64 channels per accumulator
                    umma_tf32(tmem + TC_NCHUNK * TC_N, da_lo, db_hi, (ch | ks) != 0);          // corrections share one accumulator
                    umma_tf32(tmem + TC_NCHUNK * TC_N, da_hi, db_lo, 1);
                }
                umma_commit(smem_u32(&sh->mma_done[stage]));
                if (ch == TC_NCHUNK - 1) umma_commit(smem_u32(&sh->acc_full));
            }
#pragma unroll
            for (int u = 0; u < TC_U; ++u) cur[u] = nxt[u];
        }
        xp = xp_next;
        // ---- epilogue: lane = TMEM lane = pixel ----
        if (cq == 0) {
            mbar_wait(smem_u32(&sh->acc_full), tile_it & 1u);
            tc_fence_after();
            float val[NP * 8];
            const uint32_t t_lane = tmem + ((uint32_t)(pg * 32) << 16);
#pragma unroll
            for (int part = 0; part < NP; ++part) {             // pairwise, to keep few accumulator registers live
                float a0[8], a1[8], s01[8];
                tmem_ld8(t_lane + 0 * TC_N + part * 8, a0);
                tmem_ld8(t_lane + 1 * TC_N + part * 8, a1);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i) s01[i] = a0[i] + a1[i];
                tmem_ld8(t_lane + 2 * TC_N + part * 8, a0);
                tmem_ld8(t_lane + 3 * TC_N + part * 8, a1);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i) s01[i] += a0[i] + a1[i];
                tmem_ld8(t_lane + 4 * TC_N + part * 8, a0);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i) val[part * 8 + i] = s01[i] + a0[i];
            }
            tc_fence_before();
            if (pix < N) {
                const float nrm_raw = sqrtf((sh->n2[0][m] + sh->n2[1][m]) + (sh->n2[2][m] + sh->n2[3][m]));
                if (ROWS) norms[pix] = nrm_raw;
                const float inv = mode == 2 ? 1.f : __frcp_rn(fmaxf(nrm_raw, 1e-12f));      // mode 2 (diagnostic): raw x . p_hat
#pragma unroll
                for (int c = 0; c < NP * 8; ++c) val[c] *= inv;
                if (mode == CSS_SIM_SOFTMAX) {            // softmax_c(cos_c / temp), evaluated in base 2
                    float mx = -INFINITY;
#pragma unroll
                    for (int c = 0; c < NP * 8; ++c)
                        if (c < C) mx = fmaxf(mx, val[c]);
                    const float k2 = 1.4426950408889634f / temp;
                    float sum = 0.f;
#pragma unroll
                    for (int c = 0; c < NP * 8; ++c) {
                        val[c] = (c < C) ? exp2f((val[c] - mx) * k2) : 0.f;
                        sum += val[c];
                    }
                    const float inv_sum = __frcp_rn(sum);
#pragma unroll
                    for (int c = 0; c < NP * 8; ++c) val[c] *= inv_sum;
                }
                const int b = pix / hw, s = pix - b * hw;
                float* o = out + (size_t)b * C * hw + s;
#pragma unroll
                for (int c = 0; c < NP * 8; ++c)
                    if (c < C) o[(size_t)c * hw] = val[c];
            }
        }
    }
    // ---- teardown: every MMA has completed (the last acc_full was waited for), release tensor memory ----
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TC_TMEM_COLS));
    }
}

// host side: returns 0 if the tensor-core pass was launched, <0 / >0 on error
int css_rep_pass_tc(const float* rep, const float* prototypes, float* proto_scratch, int B, int C, int h, int w, int mode, float temp,
                    float* sim_out, float* rows, float* norms, cudaStream_t st) {
    const int hw = h * w, N = B * hw;
    char* image = reinterpret_cast<char*>(proto_scratch);
    proto_prep_tc_kernel<<<TC_N, CSS_D, 0, st>>>(prototypes, image, C);
    const int n_tiles = (N + TC_M - 1) / TC_M;
    const int sms = css_cached_sm_count();
    const int grid = n_tiles < sms ? n_tiles : sms;
    cudaError_t e;
#define TC_LAUNCH(ROWS_, NP_)                                                                                                         \
    do {                                                                                                                              \
        e = cudaFuncSetAttribute(rep_pass_tc_kernel<ROWS_, NP_>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);          \
        if (e == cudaSuccess)                                                                                                         \
            rep_pass_tc_kernel<ROWS_, NP_><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(rep, image, hw, N, C, mode, temp, sim_out, rows, norms); \
    } while (0)
    if (rows) {
        if (C <= 24) TC_LAUNCH(true, 3);
        else TC_LAUNCH(true, 4);
    } else {
        if (C <= 24) TC_LAUNCH(false, 3);
        else TC_LAUNCH(false, 4);
    }
#undef TC_LAUNCH
    if (e != cudaSuccess) {
        css_set_error("css_rep_pass (tensor-core path): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}
