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
// Shape of one CTA (704 threads, one per SM, persistent over tiles of <= 128 consecutive pixels of ONE image), warp-specialised,
// mbarrier-paced, no CTA-wide barrier in the loop:
//   * loader warp: the map comes in with 1-D bulk copies (cp.async.bulk global -> shared, completion on an mbarrier).  The NCHW
//     planes are only 4-byte aligned (h*w is odd), a bulk copy needs 16-byte alignment, so lane d copies the 16-byte-aligned
//     WINDOW that covers the tile's 128 pixels of plane d (<= 528 bytes) and the readers skip the 0 / 4 / 8 / 12 bytes of lead-in.
//     Four 32-channel chunks (4 x 16.5 KB) are in flight per CTA, tracked by the barrier's transaction count -- not by registers
//     and scoreboards (a register ring of the same depth ran at one memory latency per chunk) and not by 4-byte cp.async (LDGSTS
//     issue alone took longer than the FFMA2 kernel); 512-byte bursts per plane instead of 128-byte warp requests;
//   * 16 producer warps, warp w = (pixel group w & 3, channel octet w >> 2): lane = one pixel, 8 consecutive channels of the chunk:
//     reads its own 8 raw values, splits them into TF32 hi / lo, stores them K-major SWIZZLE_128B (row = pixel, 128 B = 32
//     channels: two 128-bit stores per half; the XOR swizzle spreads the 8 pixels of a store phase over all banks) and writes
//     its 32 bytes of the pixel-major row as one whole sector;
//     (an MN-major A descriptor -- pixels contiguous, as the NCHW planes are -- returned zeros for kind::tf32 in every
//     canonical layout tried on B200, tools/dev/dev_umma.cu; K-major is also the cheaper store pattern)
//   * two A operand stages (hi | lo, 32 KB each); B (32 class slots x 256 channels, K-major SWIZZLE_128B, hi | lo = 64 KB) is
//     written once per call in exactly the shared-memory image by proto_prep_tc_kernel and copied in at kernel start;
//   * one MMA warp: waits for a stage to be full, issues its 12 MMAs (M = 128, N = 32, K = 8), commits them to the stage's
//     "empty" barrier; the accumulators are double-buffered in TMEM (2 x 160 columns), so the next tile's MMAs run while
//   * four epilogue warps (lane = TMEM lane = pixel) read the previous tile: tcgen05.ld of the five accumulators,
//     1 / max(||x||, 1e-12), optional softmax, coalesced NCHW stores, norms.
#include <stdlib.h>
#include <cuda.h>          // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, libcuda is not linked)

#include "css_common.cuh"

#define TC_PRODUCERS 512               // 16 producer warps
#define TC_THREADS (TC_PRODUCERS + 4 * 32 + 32 + 32)    // + 4 epilogue warps + MMA warp + loader warp
#define TC_M 128                      // pixels per tile  (UMMA M)
#define TC_N 32                       // class slots      (UMMA N)
#define TC_KC 32                      // channels per chunk (= per shared-memory stage = one 128-byte swizzle row)
#define TC_NCHUNK (CSS_D / TC_KC)     // 8 chunks per tile
#define TC_NACC 4                     // main accumulators per tile (two chunks = 64 channels = 8 MMAs each)
#define TC_U 8                        // channels per producer thread per chunk
#define TC_STAGES 2                   // K-major hi | lo operand stages (the MMAs drain a stage in a few hundred cycles)
#define TC_RAW 4                      // raw fp32 chunks in flight per CTA
#define TC_RAW_PITCH (TC_M * 4 + 16)  // bytes per plane window: 128 pixels + up to 12 bytes of lead-in, rounded to 16
#define TC_RAW_BYTES (TC_KC * TC_RAW_PITCH)       // 16.5 KB per raw chunk
#define TC_STAGE_HALF (TC_KC * TC_M * 4)          // 16 KB: hi (or lo) part of one stage
#define TC_B_HALF (TC_N * CSS_D * 4)              // 32 KB: hi (or lo) image of the prototypes
#define TC_ACC_COLS ((TC_NACC + 1) * TC_N)        // 160 columns per accumulator buffer
#define TC_TMEM_COLS 512                          // two buffers, rounded up to a power of two
#define TC_SMEM_BYTES (2 * TC_B_HALF + TC_STAGES * 2 * TC_STAGE_HALF + TC_RAW * TC_RAW_BYTES + 4096 + 1024)   // B | stages | raw ring | state | align

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
    css_pdl_enter();
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
    unsigned long long raw_full[TC_RAW];   // loader -> producers: the raw chunk has landed     (1 arrival + transaction bytes)
    unsigned long long raw_empty[TC_RAW];  // producers -> loader: the raw chunk has been read  (16 warp arrivals)
    unsigned long long full[TC_STAGES];    // producers -> MMA warp: the stage holds a chunk    (16 warp arrivals)
    unsigned long long empty[TC_STAGES];   // MMA warp -> producers: the MMAs that read the stage are done (tcgen05.commit)
    unsigned long long acc_full[2];        // MMA warp -> epilogue: the tile's accumulators are complete (tcgen05.commit)
    unsigned long long acc_empty[2];       // epilogue -> MMA warp: the buffer has been read               (4 warp arrivals)
    unsigned long long n2_full[2];         // producers -> epilogue: the tile's norm partials are written  (16 warp arrivals)
    uint32_t tmem_base;
    uint32_t pad;
    float n2[2][4][TC_M];                  // [tile parity][channel octet of the chunk][pixel] partial ||x||^2
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(bar) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------------------
// Tile t = (image b = t / tpi, pixels s0 = (t % tpi) * 128 .. of that image): a tile never straddles two images, so the 128
// pixels of a plane are one contiguous run.  NP: 8-column parts of the accumulators the epilogue reads (3 when C <= 24, else 4).
template <bool ROWS, int NP>
__global__ void __launch_bounds__(TC_THREADS, 1) rep_pass_tc_kernel(const float* __restrict__ rep, const char* __restrict__ b_image, int hw,
                                                                    int n_img, int C, int mode, float temp, float* __restrict__ out,
                                                                    float* __restrict__ rows, float* __restrict__ norms) {
    css_pdl_enter();
    extern __shared__ char smem_raw[];
    char* smem = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B atoms: 1 KB aligned
    char* sB = smem;                                        // [hi 32 KB | lo 32 KB]
    char* sA = smem + 2 * TC_B_HALF;                        // stage s at s * 32 KB: [hi 16 KB | lo 16 KB]
    char* sR = sA + TC_STAGES * 2 * TC_STAGE_HALF;         // raw ring: slot r at r * 16.5 KB, plane window d at d * TC_RAW_PITCH
    TcShared* sh = reinterpret_cast<TcShared*>(sR + TC_RAW * TC_RAW_BYTES);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int MMA_WARP = TC_PRODUCERS / 32 + 4, LOAD_WARP = MMA_WARP + 1;

    // ---- one-time set-up: barriers, tensor memory, the prototype image ----
    if (tid == 0) {
        for (int i = 0; i < TC_RAW; ++i) {
            mbar_init(smem_u32(&sh->raw_full[i]), 1);
            mbar_init(smem_u32(&sh->raw_empty[i]), TC_PRODUCERS / 32);
        }
        for (int i = 0; i < TC_STAGES; ++i) {
            mbar_init(smem_u32(&sh->full[i]), TC_PRODUCERS / 32);
            mbar_init(smem_u32(&sh->empty[i]), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&sh->acc_full[i]), 1);
            mbar_init(smem_u32(&sh->acc_empty[i]), 4);
            mbar_init(smem_u32(&sh->n2_full[i]), TC_PRODUCERS / 32);
        }
        fence_barrier_init();
    }
    if (warp == MMA_WARP) {
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
    const int tpi = (hw + TC_M - 1) / TC_M;                 // tiles per image
    const int n_tiles = n_img * tpi;
    const uintptr_t rep_begin = reinterpret_cast<uintptr_t>(rep), rep_end = rep_begin + (size_t)n_img * CSS_D * hw * 4;

    if (warp == LOAD_WARP) {
        // =========================================== loader ===========================================
        uint32_t g = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int b = tile / tpi, s0 = (tile - b * tpi) * TC_M, len = min(TC_M, hw - s0);
            for (int ch = 0; ch < TC_NCHUNK; ++ch, ++g) {
                const uint32_t rslot = g & (TC_RAW - 1), use = g / TC_RAW;
                if (use > 0) mbar_wait(smem_u32(&sh->raw_empty[rslot]), (use - 1u) & 1u);
                // lane d: the aligned window of plane ch*32 + d that covers pixels s0 .. s0 + len
                const uintptr_t src = rep_begin + ((size_t)(b * CSS_D + ch * TC_KC + lane) * hw + s0) * 4;
                const uintptr_t src_al = src & ~(uintptr_t)15;
                uint32_t bytes = (uint32_t)(((src - src_al) + (size_t)len * 4 + 15) & ~(size_t)15);
                const uint32_t dst = smem_u32(sR) + rslot * TC_RAW_BYTES + lane * TC_RAW_PITCH;
                const uint32_t bar = smem_u32(&sh->raw_full[rslot]);
                const bool inside = src_al >= rep_begin && src_al + bytes <= rep_end;       // never read outside the tensor
                if (!inside) {                                   // (at most the first and the last plane of the whole map)
                    float* d = reinterpret_cast<float*>(sR + rslot * TC_RAW_BYTES + lane * TC_RAW_PITCH + (src - src_al));
                    const float* sp = reinterpret_cast<const float*>(src);
                    for (int i = 0; i < len; ++i) d[i] = sp[i];
                    bytes = 0;
                }
                uint32_t total = bytes;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
                __syncwarp();
                if (lane == 0) mbar_arrive_expect_tx(bar, total);
                __syncwarp();
                if (inside) bulk_g2s(dst, reinterpret_cast<const void*>(src_al), bytes, bar);
            }
        }
    } else if (warp < TC_PRODUCERS / 32) {
        // =========================================== producers ===========================================
        const int pg = warp & 3, co = warp >> 2;            // pixel group, channel octet of every chunk
        const int m = pg * 32 + lane;                       // pixel of the tile = row of A
        // this thread's pixel row inside a stage half (K-major SWIZZLE_128B, 8-row atoms of 1 KB); its 8 channels are the
        // 16-byte chunks 2*co, 2*co + 1 of the row, XOR-swizzled by m & 7
        const uint32_t a_row = (uint32_t)((m >> 3) * 1024 + (m & 7) * 128);
        const uint32_t off0 = (uint32_t)(((2 * co) ^ (m & 7)) << 4), off1 = (uint32_t)(((2 * co + 1) ^ (m & 7)) << 4);
        const int hw3 = hw & 3, base16 = (int)(rep_begin & 15);
        uint32_t g = 0;                                     // chunks staged so far by this CTA
        uint32_t tile_it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_it) {
            const int b = tile / tpi, s0 = (tile - b * tpi) * TC_M, len = min(TC_M, hw - s0);
            const bool live = m < len;
            const int pix = b * hw + s0 + m;
            // lead-in of the plane windows: plane index mod 4 == u mod 4 for this thread's channel u (b*256, ch*32, co*8 are
            // multiples of 4), so the misalignment of its 8 windows takes 4 values per tile
            int lead[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) lead[q] = (base16 + 4 * ((q * hw3 + s0) & 3)) & 15;
            float n2 = 0.f;
#pragma unroll
            for (int ch = 0; ch < TC_NCHUNK; ++ch, ++g) {
                const int rslot = ch & (TC_RAW - 1);        // == g % TC_RAW: TC_NCHUNK is a multiple of TC_RAW
                const uint32_t slot = g & (TC_STAGES - 1), use = g / TC_STAGES;
                mbar_wait(smem_u32(&sh->raw_full[rslot]), (g / TC_RAW) & 1u);
                float x[TC_U];
                {
                    const char* rp = sR + rslot * TC_RAW_BYTES + (co * TC_U) * TC_RAW_PITCH + m * 4;
#pragma unroll
                    for (int u = 0; u < TC_U; ++u)
                        x[u] = live ? *reinterpret_cast<const float*>(rp + u * TC_RAW_PITCH + lead[u & 3]) : 0.f;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&sh->raw_empty[rslot]));       // the loader may refill the raw slot
                // pixel-major rows: the thread's 8 channels are one whole 32-byte sector of its pixel's row
                if (ROWS && live) {
                    float4* r = reinterpret_cast<float4*>(rows + (size_t)pix * CSS_D + ch * TC_KC + co * TC_U);
                    r[0] = make_float4(x[0], x[1], x[2], x[3]);
                    r[1] = make_float4(x[4], x[5], x[6], x[7]);
                }
                uint32_t hi[TC_U], lo[TC_U];
#pragma unroll
                for (int u = 0; u < TC_U; ++u) {
                    hi[u] = rna_tf32(x[u]);
                    lo[u] = rna_tf32(__fsub_rn(x[u], __uint_as_float(hi[u])));
                    n2 = fmaf(x[u], x[u], n2);
                }
                // the operand stage is free once the MMAs of its previous use have completed
                if (use > 0) mbar_wait(smem_u32(&sh->empty[slot]), (use - 1u) & 1u);
                char* st_hi = sA + slot * (2 * TC_STAGE_HALF) + a_row;
                char* st_lo = st_hi + TC_STAGE_HALF;
                *reinterpret_cast<uint4*>(st_hi + off0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(st_hi + off1) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
                *reinterpret_cast<uint4*>(st_lo + off0) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                *reinterpret_cast<uint4*>(st_lo + off1) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
                if (ch == TC_NCHUNK - 1) sh->n2[tile_it & 1u][co][m] = n2;
                fence_proxy_async();                        // generic-proxy stores -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(smem_u32(&sh->full[slot]));
                    if (ch == TC_NCHUNK - 1) mbar_arrive(smem_u32(&sh->n2_full[tile_it & 1u]));
                }
            }
        }
    } else if (warp == MMA_WARP) {
        // =========================================== MMA issuer ===========================================
        const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
        uint32_t g = 0;
        uint32_t tile_it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_it) {
            const uint32_t buf = tile_it & 1u, d_base = tmem + buf * TC_ACC_COLS;
            if (tile_it >= 2) mbar_wait(smem_u32(&sh->acc_empty[buf]), ((tile_it >> 1) - 1u) & 1u);
            tc_fence_after();
            for (int ch = 0; ch < TC_NCHUNK; ++ch, ++g) {
                const uint32_t slot = g & (TC_STAGES - 1), use = g / TC_STAGES;
                mbar_wait(smem_u32(&sh->full[slot]), use & 1u);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_hi = sA_u + slot * (2 * TC_STAGE_HALF), a_lo = a_hi + TC_STAGE_HALF;
                    const uint32_t b_blk = (uint32_t)ch * 4096;                                  // 32-channel block of the prototype image
#pragma unroll
                    for (int ks = 0; ks < TC_KC / 8; ++ks) {
                        const uint64_t da_hi = umma_desc(a_hi + ks * 32, 16, 1024), da_lo = umma_desc(a_lo + ks * 32, 16, 1024);
                        const uint64_t db_hi = umma_desc(sB_u + b_blk + ks * 32, 16, 1024), db_lo = umma_desc(sB_u + TC_B_HALF + b_blk + ks * 32, 16, 1024);
                        umma_tf32(d_base + (ch >> 1) * TC_N, da_hi, db_hi, ((ch & 1) | ks) != 0);     // main: 64 channels per accumulator
                        umma_tf32(d_base + TC_NACC * TC_N, da_lo, db_hi, (ch | ks) != 0);              // corrections share one accumulator
                        umma_tf32(d_base + TC_NACC * TC_N, da_hi, db_lo, 1);
                    }
                    umma_commit(smem_u32(&sh->empty[slot]));
                    if (ch == TC_NCHUNK - 1) umma_commit(smem_u32(&sh->acc_full[buf]));
                }
                __syncwarp();
            }
        }
    } else {
        // =========================================== epilogue ===========================================
        const int pg = warp & 3;                            // warps 16..19: TMEM lane quarter = warp % 4
        const int m = pg * 32 + lane;
        uint32_t tile_it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_it) {
            const uint32_t buf = tile_it & 1u, par = (tile_it >> 1) & 1u;
            const int b = tile / tpi, s0 = (tile - b * tpi) * TC_M, len = min(TC_M, hw - s0);
            const int s = s0 + m, pix = b * hw + s;
            mbar_wait(smem_u32(&sh->n2_full[buf]), par);
            const float* n2p = &sh->n2[buf][0][m];
            const float n2 = (n2p[0] + n2p[TC_M]) + (n2p[2 * TC_M] + n2p[3 * TC_M]);
            mbar_wait(smem_u32(&sh->acc_full[buf]), par);
            tc_fence_after();
            float val[NP * 8];
            const uint32_t t_lane = tmem + buf * TC_ACC_COLS + ((uint32_t)(pg * 32) << 16);
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
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&sh->acc_empty[buf]));      // the MMA warp may overwrite this buffer
            if (m < len) {
                const float nrm_raw = sqrtf(n2);
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
                float* o = out + (size_t)b * C * hw + s;
#pragma unroll
                for (int c = 0; c < NP * 8; ++c)
                    if (c < C) o[(size_t)c * hw] = val[c];
            }
        }
    }
    // ---- teardown: all roles are done (the epilogue waited for the last accumulators), release tensor memory ----
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TC_TMEM_COLS));
    }
}

// host side: returns 0 if the tensor-core pass was launched, <0 / >0 on error
int css_rep_pass_tc(const float* rep, const float* prototypes, float* proto_scratch, int B, int C, int h, int w, int mode, float temp,
                    float* sim_out, float* rows, float* norms, cudaStream_t st) {
    const int hw = h * w;
    char* image = reinterpret_cast<char*>(proto_scratch);
    css_launch(proto_prep_tc_kernel, dim3(TC_N), dim3(CSS_D), (size_t)(0), (cudaStream_t)(st), prototypes, image, C);
    const int n_tiles = B * ((hw + TC_M - 1) / TC_M);
    const int sms = css_cached_sm_count();
    const int grid = n_tiles < sms ? n_tiles : sms;
    cudaError_t e;
#define TC_LAUNCH(ROWS_, NP_)                                                                                                         \
    do {                                                                                                                              \
        e = cudaFuncSetAttribute(rep_pass_tc_kernel<ROWS_, NP_>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);          \
        if (e == cudaSuccess)                                                                                                         \
            css_launch(rep_pass_tc_kernel<ROWS_, NP_>, dim3(grid), dim3(TC_THREADS), (size_t)(TC_SMEM_BYTES), (cudaStream_t)(st), rep, image, hw, B, C, mode, temp, sim_out, rows, norms); \
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


// =====================================================================================================================
// Channels-last maps (torch.channels_last: memory is [pixel][256], i.e. the map IS the pixel-major row table of the loss)
// =====================================================================================================================
// With rows contiguous and 1 KB aligned the pass needs no transposition and no copy at all:
//   * one thread issues a 2-D TMA tensor copy per 32-channel chunk ([128 rows x 32 channels] box, SWIZZLE_128B): the tile lands in
//     shared memory in exactly the K-major layout the tensor core reads, NL_RAW chunks in flight per CTA;
//   * the tensor core TRUNCATES an fp32 bit pattern to TF32 (tools/dev/dev_umma.cu), so the raw tile is the "hi" operand as it
//     is; 16 producer warps only compute lo = x - trunc_tf32(x) (exact) into a second tile at the same swizzled positions, and
//     the norm partials;
//   * MMA warp / TMEM accumulators / epilogue as in the NCHW kernel above.  Output: sim / prob [B,C,h,w] and ||x_p||.
#define NL_RAW 6                                  // raw chunks in flight per CTA (16 KB each)
#define NL_LO 2                                   // lo stages
#define NL_SMEM_BYTES (2 * TC_B_HALF + NL_RAW * TC_STAGE_HALF + NL_LO * TC_STAGE_HALF + 4096 + 1024)

struct NlShared {
    unsigned long long raw_full[NL_RAW];    // TMA -> producers, MMA warp: the raw chunk has landed          (1 arrival + transaction bytes)
    unsigned long long raw_empty[NL_RAW];   // producers (16 warp arrivals) + MMA warp (tcgen05.commit) -> TMA thread: slot is free
    unsigned long long lo_full[NL_LO];      // producers -> MMA warp: the lo tile is written                  (16 warp arrivals)
    unsigned long long lo_empty[NL_LO];     // MMA warp -> producers: the MMAs that read the lo tile are done (tcgen05.commit)
    unsigned long long acc_full[2];
    unsigned long long acc_empty[2];
    unsigned long long n2_full[2];
    uint32_t tmem_base;
    uint32_t pad;
    float n2[2][4][TC_M];
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tmap, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(tmap), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

template <int NP>
__global__ void __launch_bounds__(TC_THREADS, 1) rep_pass_nhwc_kernel(const __grid_constant__ CUtensorMap tmap, const char* __restrict__ b_image,
                                                                      int hw, int N, int C, int mode, float temp, float* __restrict__ out,
                                                                      float* __restrict__ norms, const int32_t* __restrict__ guard, int dbg) {
    css_pdl_enter();
    if (guard != nullptr && *guard == 0) return;            // css_rows_refresh_nhwc: the carried norms were verified, nothing to redo
    extern __shared__ char smem_raw[];
    char* smem = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    char* sB = smem;                                        // [hi 32 KB | lo 32 KB]
    char* sR = smem + 2 * TC_B_HALF;                        // raw ring: slot r at r * 16 KB, K-major SWIZZLE_128B (written by TMA)
    char* sL = sR + NL_RAW * TC_STAGE_HALF;                 // lo stages
    NlShared* sh = reinterpret_cast<NlShared*>(sL + NL_LO * TC_STAGE_HALF);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int MMA_WARP = TC_PRODUCERS / 32 + 4, LOAD_WARP = MMA_WARP + 1;

    if (tid == 0) {
        for (int i = 0; i < NL_RAW; ++i) {
            mbar_init(smem_u32(&sh->raw_full[i]), 1);
            mbar_init(smem_u32(&sh->raw_empty[i]), TC_PRODUCERS / 32 + 1);
        }
        for (int i = 0; i < NL_LO; ++i) {
            mbar_init(smem_u32(&sh->lo_full[i]), TC_PRODUCERS / 32);
            mbar_init(smem_u32(&sh->lo_empty[i]), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&sh->acc_full[i]), 1);
            mbar_init(smem_u32(&sh->acc_empty[i]), 4);
            mbar_init(smem_u32(&sh->n2_full[i]), TC_PRODUCERS / 32);
        }
        fence_barrier_init();
    }
    if (warp == MMA_WARP) {
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
    const int n_tiles = (N + TC_M - 1) / TC_M;

    if (warp == LOAD_WARP) {
        // =========================================== TMA issuer ===========================================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
            uint32_t g = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int ch = 0; ch < TC_NCHUNK; ++ch, ++g) {
                    const uint32_t rslot = g % NL_RAW, use = g / NL_RAW;
                    if (use > 0) mbar_wait(smem_u32(&sh->raw_empty[rslot]), (use - 1u) & 1u);
                    const uint32_t bar = smem_u32(&sh->raw_full[rslot]);
                    mbar_arrive_expect_tx(bar, TC_STAGE_HALF);                   // rows past the end of the map are zero-filled
                    tma_load_2d(smem_u32(sR) + rslot * TC_STAGE_HALF, &tmap, ch * TC_KC, tile * TC_M, bar);
                }
            }
        }
    } else if (warp < TC_PRODUCERS / 32) {
        // =========================================== producers: lo tile + norm partials ===========================================
        const int pg = warp & 3, co = warp >> 2;
        const int m = pg * 32 + lane;
        const uint32_t a_row = (uint32_t)((m >> 3) * 1024 + (m & 7) * 128);
        const uint32_t off0 = (uint32_t)(((2 * co) ^ (m & 7)) << 4), off1 = (uint32_t)(((2 * co + 1) ^ (m & 7)) << 4);
        uint32_t g = 0, tile_it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_it) {
            float n2 = 0.f;
            for (int ch = 0; ch < TC_NCHUNK; ++ch, ++g) {
                const uint32_t rslot = g % NL_RAW, slot = g % NL_LO, use = g / NL_LO;
                mbar_wait(smem_u32(&sh->raw_full[rslot]), (g / NL_RAW) & 1u);
                const char* rp = sR + rslot * TC_STAGE_HALF + a_row;
                const uint4 q0 = *reinterpret_cast<const uint4*>(rp + off0), q1 = *reinterpret_cast<const uint4*>(rp + off1);
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&sh->raw_empty[rslot]));       // (the MMA warp's commit is the 17th arrival)
                const uint32_t xs[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
                uint32_t lo[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float x = __uint_as_float(xs[u]);
                    lo[u] = __float_as_uint(__fsub_rn(x, __uint_as_float(xs[u] & 0xFFFFE000u)));      // exact: x - trunc_tf32(x)
                    n2 = fmaf(x, x, n2);
                }
                if (use > 0) mbar_wait(smem_u32(&sh->lo_empty[slot]), (use - 1u) & 1u);
                char* lp = sL + slot * TC_STAGE_HALF + a_row;
                *reinterpret_cast<uint4*>(lp + off0) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                *reinterpret_cast<uint4*>(lp + off1) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
                if (ch == TC_NCHUNK - 1) sh->n2[tile_it & 1u][co][m] = n2;
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(smem_u32(&sh->lo_full[slot]));
                    if (ch == TC_NCHUNK - 1) mbar_arrive(smem_u32(&sh->n2_full[tile_it & 1u]));
                }
            }
        }
    } else if (warp == MMA_WARP) {
        // =========================================== MMA issuer ===========================================
        const uint32_t sR_u = smem_u32(sR), sL_u = smem_u32(sL), sB_u = smem_u32(sB);
        uint32_t g = 0, tile_it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_it) {
            const uint32_t buf = tile_it & 1u, d_base = tmem + buf * TC_ACC_COLS;
            if (tile_it >= 2) mbar_wait(smem_u32(&sh->acc_empty[buf]), ((tile_it >> 1) - 1u) & 1u);
            tc_fence_after();
            for (int ch = 0; ch < TC_NCHUNK; ++ch, ++g) {
                const uint32_t rslot = g % NL_RAW, slot = g % NL_LO;
                mbar_wait(smem_u32(&sh->raw_full[rslot]), (g / NL_RAW) & 1u);
                mbar_wait(smem_u32(&sh->lo_full[slot]), (g / NL_LO) & 1u);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_hi = sR_u + rslot * TC_STAGE_HALF, a_lo = sL_u + slot * TC_STAGE_HALF;
                    const uint32_t b_blk = (uint32_t)ch * 4096;
#pragma unroll
                    for (int ks = 0; ks < TC_KC / 8; ++ks) {
                        const uint64_t da_hi = umma_desc(a_hi + ks * 32, 16, 1024), da_lo = umma_desc(a_lo + ks * 32, 16, 1024);
                        const uint64_t db_hi = umma_desc(sB_u + b_blk + ks * 32, 16, 1024), db_lo = umma_desc(sB_u + TC_B_HALF + b_blk + ks * 32, 16, 1024);
                        if (!(dbg & 1)) umma_tf32(d_base + (ch >> 1) * TC_N, da_hi, db_hi, ((ch & 1) | ks) != 0);
                        if (!(dbg & 3)) umma_tf32(d_base + TC_NACC * TC_N, da_lo, db_hi, (ch | ks) != 0);
                        if (!(dbg & 3)) umma_tf32(d_base + TC_NACC * TC_N, da_hi, db_lo, 1);
                    }
                    umma_commit(smem_u32(&sh->raw_empty[rslot]));
                    umma_commit(smem_u32(&sh->lo_empty[slot]));
                    if (ch == TC_NCHUNK - 1) umma_commit(smem_u32(&sh->acc_full[buf]));
                }
                __syncwarp();
            }
        }
    } else {
        // =========================================== epilogue ===========================================
        const int pg = warp & 3;
        const int m = pg * 32 + lane;
        uint32_t tile_it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_it) {
            const uint32_t buf = tile_it & 1u, par = (tile_it >> 1) & 1u;
            const int pix = tile * TC_M + m;
            mbar_wait(smem_u32(&sh->n2_full[buf]), par);
            const float* n2p = &sh->n2[buf][0][m];
            const float n2 = (n2p[0] + n2p[TC_M]) + (n2p[2 * TC_M] + n2p[3 * TC_M]);
            mbar_wait(smem_u32(&sh->acc_full[buf]), par);
            tc_fence_after();
            float val[NP * 8];
            const uint32_t t_lane = tmem + buf * TC_ACC_COLS + ((uint32_t)(pg * 32) << 16);
#pragma unroll
            for (int part = 0; part < NP; ++part) {
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
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&sh->acc_empty[buf]));
            if (pix < N) {
                const float nrm_raw = sqrtf(n2);
                if (norms) norms[pix] = nrm_raw;
                if (out) {
                    const float inv = __frcp_rn(fmaxf(nrm_raw, 1e-12f));
#pragma unroll
                    for (int c = 0; c < NP * 8; ++c) val[c] *= inv;
                    if (mode == CSS_SIM_SOFTMAX) {
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
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TC_TMEM_COLS));
    }
}

typedef CUresult (*css_tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                       const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                       CUtensorMapFloatOOBfill);

static css_tmap_encode_fn css_tmap_encoder() {
    static css_tmap_encode_fn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (css_tmap_encode_fn)p;
    }
    return fn;
}

// channels-last rep pass: rows = the map itself ([N][256] fp32, 16-byte aligned).  sim_out / norms may each be NULL.
int css_rep_pass_nhwc_tc(const float* rows, const float* prototypes, float* proto_scratch, int B, int C, int h, int w, int mode, float temp,
                         float* sim_out, float* norms, const int32_t* guard, cudaStream_t st) {
    const int hw = h * w, N = B * hw;
    css_tmap_encode_fn enc = css_tmap_encoder();
    if (!enc) {
        css_set_error("css_rep_pass: cuTensorMapEncodeTiled is not available from this driver");
        return CSS_E_ARG;
    }
    CUtensorMap tmap;
    const cuuint64_t dims[2] = {(cuuint64_t)CSS_D, (cuuint64_t)N};
    const cuuint64_t strides[1] = {(cuuint64_t)CSS_D * 4};
    const cuuint32_t box[2] = {TC_KC, TC_M};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(rows), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        css_set_error("css_rep_pass: cuTensorMapEncodeTiled failed (%d)", (int)r);
        return CSS_E_ARG;
    }
    char* image = reinterpret_cast<char*>(proto_scratch);
    if (prototypes) css_launch(proto_prep_tc_kernel, dim3(TC_N), dim3(CSS_D), (size_t)0, st, prototypes, image, C);
    else cudaMemsetAsync(image, 0, 2 * TC_B_HALF, st);
    const int n_tiles = (N + TC_M - 1) / TC_M;
    const int sms = css_cached_sm_count();
    const int grid = n_tiles < sms ? n_tiles : sms;
    const char* dbg_env = getenv("CSS_B200_NHWC_DBG");      // development knob: 1 = no MMAs, 2 = no correction MMAs (results are wrong)
    const int dbg = dbg_env ? atoi(dbg_env) : 0;
    cudaError_t e;
    if (C <= 24) {
        e = cudaFuncSetAttribute(rep_pass_nhwc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, NL_SMEM_BYTES);
        if (e == cudaSuccess) e = css_launch(rep_pass_nhwc_kernel<3>, dim3(grid), dim3(TC_THREADS), (size_t)NL_SMEM_BYTES, st, tmap, (const char*)image, hw, N, C, mode, temp, sim_out, norms, guard, dbg);
    } else {
        e = cudaFuncSetAttribute(rep_pass_nhwc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, NL_SMEM_BYTES);
        if (e == cudaSuccess) e = css_launch(rep_pass_nhwc_kernel<4>, dim3(grid), dim3(TC_THREADS), (size_t)NL_SMEM_BYTES, st, tmap, (const char*)image, hw, N, C, mode, temp, sim_out, norms, guard, dbg);
    }
    if (e != cudaSuccess) {
        css_set_error("css_rep_pass (channels-last path): %s", cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}


extern "C" int css_rep_pass_nhwc(const void* rep_rows, int rep_dtype, const float* prototypes, float* proto_scratch, int B, int C, int D, int h,
                                 int w, int mode, float temp, float* sim_out, float* norms, void* stream) {
    CSS_CHECK_ARG(rep_rows && proto_scratch && (sim_out || norms), CSS_E_ARG, "css_rep_pass_nhwc: null pointer / nothing to do");
    CSS_CHECK_ARG(!sim_out || prototypes, CSS_E_ARG, "css_rep_pass_nhwc: sim_out needs prototypes");
    CSS_CHECK_ARG(B > 0 && h > 0 && w > 0, CSS_E_ARG, "css_rep_pass_nhwc: non-positive size");
    CSS_CHECK_ARG(mode == CSS_SIM_COS || mode == CSS_SIM_SOFTMAX, CSS_E_ARG, "css_rep_pass_nhwc: bad mode %d", mode);
    if (int e = css_check_dims(sim_out ? C : 1, D)) return e;
    CSS_CHECK_ARG(rep_dtype == CSS_DTYPE_F32, CSS_E_DTYPE, "css_rep_pass_nhwc: only float32 channels-last maps are supported (dtype %d)", rep_dtype);
    CSS_CHECK_ARG(((uintptr_t)rep_rows & 15) == 0, CSS_E_ARG, "css_rep_pass_nhwc: the map must be 16-byte aligned");
    CSS_CHECK_ARG((long long)B * h * w < (1ll << 31) / CSS_CMAX, CSS_E_SIZE, "css_rep_pass_nhwc: too many pixels");
    if (int e = css_rep_pass_nhwc_tc((const float*)rep_rows, prototypes, proto_scratch, B, sim_out ? C : 1, h, w, mode, temp, sim_out, norms, nullptr,
                                     (cudaStream_t)stream))
        return e;
    CSS_CHECK_LAUNCH("css_rep_pass_nhwc", 2);
    return 0;
}

// channels-last twin of css_rows_refresh: are the carried norms still those of THIS map?  `src_rows` is the map the norms were
// computed from (kept alive by the caller), `rep_rows` the one the loss received (e.g. DDP's clone of it).
__global__ void __launch_bounds__(256) rows_verify_rm_kernel(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, int N,
                                                             int32_t* __restrict__ meta) {
    css_pdl_enter();
    const int p = blockIdx.x * 256 + threadIdx.x;
    if (p >= N) return;
    const int d0 = (int)(((unsigned)p * 37u) & 63u);
    bool bad = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) bad |= a[(size_t)p * CSS_D + d0 + 64 * i] != b[(size_t)p * CSS_D + d0 + 64 * i];
    if (bad) meta[CSS_META_ROWS_STALE] = 1;
}

extern "C" int css_rows_refresh_nhwc(const void* rep_rows, const void* src_rows, float* norms, float* proto_scratch, int32_t* meta, int B, int D,
                                     int h, int w, void* stream) {
    CSS_CHECK_ARG(rep_rows && src_rows && norms && proto_scratch && meta, CSS_E_ARG, "css_rows_refresh_nhwc: null pointer");
    CSS_CHECK_ARG(B > 0 && h > 0 && w > 0 && D == CSS_D, CSS_E_ARG, "css_rows_refresh_nhwc: bad size");
    CSS_CHECK_ARG(((uintptr_t)rep_rows & 15) == 0, CSS_E_ARG, "css_rows_refresh_nhwc: the map must be 16-byte aligned");
    const int N = B * h * w;
    cudaStream_t st = (cudaStream_t)stream;
    css_launch(rows_verify_rm_kernel, dim3((N + 255) / 256), dim3(256), (size_t)0, st, (const uint32_t*)rep_rows, (const uint32_t*)src_rows, N, meta);
    if (int e = css_rep_pass_nhwc_tc((const float*)rep_rows, nullptr, proto_scratch, B, 1, h, w, CSS_SIM_COS, 1.f, nullptr, norms,
                                     meta + CSS_META_ROWS_STALE, st))
        return e;
    CSS_CHECK_LAUNCH("css_rows_refresh_nhwc", 2);
    return 0;
}
