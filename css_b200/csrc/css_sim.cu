// Stage 1 / 1b / 1' / 2: similarity maps against the class prototypes, fused up-sample + softmax + max + fusion.
// Reference: generalframeworks/networks/ddp_model.py:104-118, 147-154 (Model_mix); :189-199, 230-237 (Model_cross);
// :36-37 (Model_ori_pseudo).
#include <stdlib.h>

#include "css_common.cuh"

#define CSS_REP_PASS_DEFAULT_TC 0
int css_rep_pass_tc(const float* rep, const float* prototypes, float* proto_scratch, int B, int C, int h, int w, int mode, float temp,
                    float* sim_out, float* rows, float* norms, cudaStream_t st);
bool css_rep_pass_use_tc();

// ---------------------------------------------------------------------------------------------------------------
// prototype preparation: F.normalize(prototypes, dim=-1) (eps 1e-12, ddp_model.py:107), transposed to [D][32]
// (class-minor, zero padded) so the streaming kernel reads 4 classes per 128-bit shared-memory load.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CSS_D) proto_prep_kernel(const float* __restrict__ protos, float* __restrict__ scratch, int C) {
    css_pdl_enter();
    __shared__ float part[CSS_D / 32];
    const int c = blockIdx.x, d = threadIdx.x;
    if (c >= C) {                                   // padding columns
        scratch[d * CSS_CMAX + c] = 0.f;
        return;
    }
    const float v = protos[c * CSS_D + d];
    const float s = warp_sum(v * v);
    if ((d & 31) == 0) part[d >> 5] = s;
    __syncthreads();
    float n2 = 0.f;
#pragma unroll
    for (int i = 0; i < CSS_D / 32; ++i) n2 += part[i];
    scratch[d * CSS_CMAX + c] = __fdiv_rn(v, fmaxf(sqrtf(n2), 1e-12f));
}

// ---------------------------------------------------------------------------------------------------------------
// K1 "rep pass": ONE streaming read of an NCHW representation map that produces any combination of
//   (a) the similarity of every pixel's 256-vector against the class prototypes
//         cos_c = (x . p_hat_c) / max(||x||, 1e-12)                             (ddp_model.py:105-109)
//         mode CSS_SIM_SOFTMAX: softmax_c(cos_c / temp)                         (ddp_model.py:154)
//   (b) the pixel-major copy rows[p][256] (raw values) + ||x_p||, from which the loss gathers 1 KB candidate rows and
//       accumulates class sums (loss.py:85,102,111-112,142: permute + boolean gathers).
// Algorithmic bytes / pixel: D*4 read (+ C*4 written) (+ D*4 + 4 written).  At 22 FMA per 4-byte element (a) sits on
// the fp32 FMA ridge of the SM (measured: FMA-issue bound, not HBM bound), so the inner loop spends as few issue slots
// as possible per element, and (b) rides along for free under the FMA time:
//   * packed fp32x2 FMAs (FFMA2, sm_100): one instruction updates two class dots;
//   * each lane owns SM_PPT pixels x half of the channels (SM_KS = 2 channel slices per warp), so one 128-bit shared
//     memory read of 4 pre-normalised prototype values feeds SM_PPT x 2 FFMA2 and the per-thread channel loop is short
//     enough to keep 32 independent 4-byte loads in flight per thread;
//   * the 16 consecutive channels a lane holds per step are 64 contiguous bytes of its pixel's row: the transpose is
//     four 128-bit stores straight from registers, no shared-memory staging;
//   * the two channel slices are combined with one xor-shuffle per value (fixed order: deterministic).
// A warp reads 2 x 64 contiguous bytes per (channel pair, pixel group); the 4 warps of a CTA cover adjacent pixels.
// ---------------------------------------------------------------------------------------------------------------
// rep element loaders: fp32, or bf16 widened exactly to fp32 (all arithmetic stays fp32)
// The map is read exactly once: its lines are marked evict-first in L2 so that they do not push out the pixel-major rows the same
// kernel is writing (class sums and the scorer read those next): -5 us per V321 step.  Marking the row stores evict-last on top of
// that was measured slower (student pass 68 -> 74 us, step +17 us).  CSS_B200_L2_HINT=0 at build time restores plain streaming loads.
#ifndef CSS_B200_L2_HINT
#define CSS_B200_L2_HINT 1
#endif
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float ld_elem(const float* p, uint64_t pol) {
#if CSS_B200_L2_HINT
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    return v;
#else
    return ldg_stream(p);
#endif
}
__device__ __forceinline__ float ld_elem(const __nv_bfloat16* p, uint64_t) {
    unsigned short u;
    asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(u) : "l"(p));
    return __uint_as_float((uint32_t)u << 16);
}

#define SM_PPT 2
#define SM_KS 2
#define SM_U 16
#define SM_WARPS 4
#define SM_MINB 4

// T1: the last class group holds ONE class (C = 4 NG - 3, e.g. VOC's 21): one scalar FMA instead of two packed ones per pixel
template <int NG, bool ROWS, typename T, bool T1 = false>
__global__ void __launch_bounds__(SM_WARPS * 32, SM_MINB) rep_pass_kernel(const T* __restrict__ rep, const float* __restrict__ scratch,
                                                                         int hw, int N, int C, int mode, float temp,
                                                                         float* __restrict__ out, T* __restrict__ rows,
                                                                         float* __restrict__ norms, const int32_t* __restrict__ guard) {
    css_pdl_enter();
    if (guard != nullptr && *guard == 0) return;     // css_rows_refresh: the carried rows were verified, nothing to redo
    constexpr int DS = CSS_D / SM_KS;          // channels per slice
    constexpr int PL = 32 / SM_KS;             // pixel lanes per warp
    constexpr int WP = PL * SM_PPT;            // pixels per warp
    constexpr int NGA = NG > 0 ? NG : 1;
    constexpr int SL = DS * NGA + 1;           // float4 per slice (+1 skew: the slices land in different banks)
    __shared__ float4 sp[NG > 0 ? SM_KS * SL : 1];
    if (NG > 0) {
        for (int i = threadIdx.x; i < CSS_D * NG; i += SM_WARPS * 32) {
            const int d = i / NGA, g = i - d * NGA;
            sp[(d / DS) * SL + (d % DS) * NGA + g] = reinterpret_cast<const float4*>(scratch)[d * (CSS_CMAX / 4) + g];
        }
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ks = lane / PL, psub = lane % PL;
    const float4* myp = sp + (NG > 0 ? ks * SL : 0);
    const uint64_t pol = l2_evict_first_policy();
    const int n_wchunks = (N + WP - 1) / WP;
    for (int wc = blockIdx.x * SM_WARPS + warp; wc < n_wchunks; wc += gridDim.x * SM_WARPS) {
        const T* x[SM_PPT];
        int pix[SM_PPT];
#pragma unroll
        for (int j = 0; j < SM_PPT; ++j) {
            pix[j] = wc * WP + j * PL + psub;
            const int p = min(pix[j], N - 1);                    // out-of-range lanes re-read the last pixel, never write
            const int b = p / hw;
            x[j] = rep + ((size_t)b * CSS_D + ks * DS) * hw + (p - b * hw);
        }
        float2 acc[SM_PPT][2 * NGA];
        float n2[SM_PPT];
#pragma unroll
        for (int j = 0; j < SM_PPT; ++j) {
            n2[j] = 0.f;
#pragma unroll
            for (int i = 0; i < 2 * NGA; ++i) acc[j][i] = make_float2(0.f, 0.f);
        }
#pragma unroll 1
        for (int d0 = 0; d0 < DS; d0 += SM_U) {
            float v[SM_U][SM_PPT];
#pragma unroll
            for (int u = 0; u < SM_U; ++u)
#pragma unroll
                for (int j = 0; j < SM_PPT; ++j) v[u][j] = ld_elem(x[j] + (size_t)(d0 + u) * hw, pol);
            if (ROWS && sizeof(T) == 2) {
                // bf16 map -> bf16 rows (lossless): the 16 channels a lane holds are 32 contiguous bytes = one whole sector
#pragma unroll
                for (int j = 0; j < SM_PPT; ++j) {
                    if (pix[j] < N) {
                        uint32_t w[SM_U / 2];
#pragma unroll
                        for (int u = 0; u < SM_U; u += 2)
                            w[u / 2] = (__float_as_uint(v[u][j]) >> 16) | (__float_as_uint(v[u + 1][j]) & 0xffff0000u);
                        uint4* r = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(rows) + (size_t)pix[j] * CSS_D + ks * DS + d0);
                        r[0] = make_uint4(w[0], w[1], w[2], w[3]);
                        r[1] = make_uint4(w[4], w[5], w[6], w[7]);
                    }
                }
            } else if (ROWS) {
                // Transposed write-out straight from registers.  A lane holds 64 contiguous bytes (16 channels) of its pixel's
                // row; neighbouring lanes (pixels p, p+1) first swap 16-byte chunks so that every 128-bit store instruction
                // completes whole 32-byte sectors (lane pair -> one sector) instead of half sectors.
                float* rows_f = reinterpret_cast<float*>(rows);
                const int odd = psub & 1;
#pragma unroll
                for (int j = 0; j < SM_PPT; ++j) {
                    const int pA = pix[j] - odd, pB = pA + 1;
                    float* rA = rows_f + (size_t)pA * CSS_D + ks * DS + d0 + 4 * odd;
                    float* rB = rA + CSS_D;
#pragma unroll
                    for (int hh = 0; hh < SM_U / 8; ++hh) {
                        const int ue = 8 * hh, uo = 8 * hh + 4;          // even / odd 16-byte chunk of this 32-byte sector
                        float4 own_e = make_float4(v[ue][j], v[ue + 1][j], v[ue + 2][j], v[ue + 3][j]);
                        float4 own_o = make_float4(v[uo][j], v[uo + 1][j], v[uo + 2][j], v[uo + 3][j]);
                        float4 snd = odd ? own_e : own_o, rcv;
                        rcv.x = __shfl_xor_sync(0xffffffffu, snd.x, 1);
                        rcv.y = __shfl_xor_sync(0xffffffffu, snd.y, 1);
                        rcv.z = __shfl_xor_sync(0xffffffffu, snd.z, 1);
                        rcv.w = __shfl_xor_sync(0xffffffffu, snd.w, 1);
                        const float4 first = odd ? rcv : own_e;          // pixel pA: even lane -> chunk e, odd lane -> chunk o
                        const float4 second = odd ? own_o : rcv;         // pixel pB
                        if (pA < N) *reinterpret_cast<float4*>(rA + 8 * hh) = first;
                        if (pB < N) *reinterpret_cast<float4*>(rB + 8 * hh) = second;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < SM_U; ++u) {
#pragma unroll
                for (int j = 0; j < SM_PPT; ++j) n2[j] = fmaf(v[u][j], v[u][j], n2[j]);
                if (NG > 0) {
#pragma unroll
                    for (int g = 0; g < NG; ++g) {
                        if (T1 && g == NG - 1) {
                            const float q1 = reinterpret_cast<const float*>(myp + (d0 + u) * NGA + g)[0];
#pragma unroll
                            for (int j = 0; j < SM_PPT; ++j) acc[j][2 * g].x = fmaf(v[u][j], q1, acc[j][2 * g].x);
                        } else {
                            const float4 q = myp[(d0 + u) * NGA + g];
#pragma unroll
                            for (int j = 0; j < SM_PPT; ++j) {
                                const float2 vv = make_float2(v[u][j], v[u][j]);
                                acc[j][2 * g + 0] = __ffma2_rn(vv, make_float2(q.x, q.y), acc[j][2 * g + 0]);
                                acc[j][2 * g + 1] = __ffma2_rn(vv, make_float2(q.z, q.w), acc[j][2 * g + 1]);
                            }
                        }
                    }
                }
            }
        }
        // combine the channel slices
#pragma unroll
        for (int o = PL; o < 32; o <<= 1) {
#pragma unroll
            for (int j = 0; j < SM_PPT; ++j) {
                n2[j] += __shfl_xor_sync(0xffffffffu, n2[j], o);
                if (NG > 0) {
#pragma unroll
                    for (int i = 0; i < 2 * NG; ++i) {
                        if (T1 && i == 2 * NG - 1) continue;                       // dead slots of the one-class tail
                        acc[j][i].x += __shfl_xor_sync(0xffffffffu, acc[j][i].x, o);
                        if (!(T1 && i == 2 * NG - 2)) acc[j][i].y += __shfl_xor_sync(0xffffffffu, acc[j][i].y, o);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < SM_PPT; ++j) {
            if (pix[j] >= N || (j % SM_KS) != ks) continue;      // slice j % KS writes pixel j
            const float nrm_raw = sqrtf(n2[j]);
            if (ROWS) norms[pix[j]] = nrm_raw;
            if (NG > 0) {
                const int b = pix[j] / hw, s = pix[j] - b * hw;
                const float inv = __frcp_rn(fmaxf(nrm_raw, 1e-12f));
                float* o = out + (size_t)b * C * hw + s;
                float val[4 * NGA];
#pragma unroll
                for (int i = 0; i < 2 * NG; ++i) {
                    val[2 * i] = acc[j][i].x * inv;
                    val[2 * i + 1] = acc[j][i].y * inv;
                }
                if (mode == CSS_SIM_SOFTMAX) {            // softmax_c(cos_c / temp), evaluated in base 2
                    float m = -INFINITY;
#pragma unroll
                    for (int c = 0; c < 4 * NG; ++c)
                        if (c < C) m = fmaxf(m, val[c]);
                    const float k2 = 1.4426950408889634f / temp;
                    float sum = 0.f;
#pragma unroll
                    for (int c = 0; c < 4 * NG; ++c) {
                        val[c] = (c < C) ? exp2f((val[c] - m) * k2) : 0.f;
                        sum += val[c];
                    }
                    const float inv_sum = __frcp_rn(sum);
#pragma unroll
                    for (int c = 0; c < 4 * NG; ++c) val[c] *= inv_sum;
                }
#pragma unroll
                for (int c = 0; c < 4 * NG; ++c)
                    if (c < C) o[(size_t)c * hw] = val[c];
            }
        }
    }
}

template <int NG, bool ROWS, typename T, bool T1 = false>
static void launch_rep_pass(const T* rep, const float* scratch, int hw, int N, int C, int mode, float temp, float* out,
                            T* rows, float* norms, cudaStream_t st, const int32_t* guard = nullptr) {
    constexpr int WP = (32 / SM_KS) * SM_PPT;
    const int n_blocks = ((N + WP - 1) / WP + SM_WARPS - 1) / SM_WARPS;
    const int cap = css_cached_sm_count() * SM_MINB;
    css_launch(rep_pass_kernel<NG, ROWS, T, T1>, dim3(n_blocks < cap ? n_blocks : cap), dim3(SM_WARPS * 32), (size_t)0, st, rep, scratch, hw, N, C, mode, temp, out,
                                                                                            rows, norms, guard);
}

template <bool ROWS, typename T>
static void dispatch_rep_pass(int ng, const T* rep, const float* scratch, int hw, int N, int C, int mode, float temp, float* out,
                              T* rows, float* norms, cudaStream_t st) {
    switch (ng) {
        case 1: launch_rep_pass<1, ROWS, T>(rep, scratch, hw, N, C, mode, temp, out, rows, norms, st); break;
        case 2: launch_rep_pass<2, ROWS, T>(rep, scratch, hw, N, C, mode, temp, out, rows, norms, st); break;
        case 3: launch_rep_pass<3, ROWS, T>(rep, scratch, hw, N, C, mode, temp, out, rows, norms, st); break;
        case 4: launch_rep_pass<4, ROWS, T>(rep, scratch, hw, N, C, mode, temp, out, rows, norms, st); break;
        case 5: launch_rep_pass<5, ROWS, T>(rep, scratch, hw, N, C, mode, temp, out, rows, norms, st); break;
        case 6:
            if (C == 21) launch_rep_pass<6, ROWS, T, true>(rep, scratch, hw, N, C, mode, temp, out, rows, norms, st);   // VOC
            else launch_rep_pass<6, ROWS, T>(rep, scratch, hw, N, C, mode, temp, out, rows, norms, st);
            break;
        case 7: launch_rep_pass<7, ROWS, T>(rep, scratch, hw, N, C, mode, temp, out, rows, norms, st); break;
        default: launch_rep_pass<8, ROWS, T>(rep, scratch, hw, N, C, mode, temp, out, rows, norms, st); break;
    }
}

// which multiply side css_rep_pass uses for fp32 maps: the tcgen05 / TMEM kernel of css_sim_tc.cu ("tc") or the FFMA2 kernel
// above ("fma"); CSS_B200_REP_PASS overrides the default, css_set_rep_pass_path() overrides both (tests time and compare the two)
static int g_rep_pass_tc = -1;
extern "C" int css_set_rep_pass_path(int use_tc) {
    g_rep_pass_tc = use_tc < 0 ? -1 : (use_tc ? 1 : 0);
    return 0;
}
bool css_rep_pass_use_tc() {
    if (g_rep_pass_tc >= 0) return g_rep_pass_tc != 0;
    static int env = -1;
    if (env < 0) {
        const char* v = getenv("CSS_B200_REP_PASS");
        env = (v && v[0] == 't') ? 1 : (v && v[0] == 'f') ? 0 : CSS_REP_PASS_DEFAULT_TC;
    }
    return env != 0;
}

extern "C" int css_rep_pass(const void* rep, int rep_dtype, const float* prototypes, float* proto_scratch, int B, int C, int D, int h,
                            int w, int mode, float temp, float* sim_out, void* rows, float* norms, void* stream) {
    const bool want_sim = sim_out != nullptr, want_rows = rows != nullptr;
    CSS_CHECK_ARG(rep && (want_sim || want_rows), CSS_E_ARG, "css_rep_pass: null pointer / nothing to do");
    CSS_CHECK_ARG(!want_sim || (prototypes && proto_scratch), CSS_E_ARG, "css_rep_pass: sim_out needs prototypes and proto_scratch");
    CSS_CHECK_ARG(want_rows == (norms != nullptr), CSS_E_ARG, "css_rep_pass: rows and norms go together");
    CSS_CHECK_ARG(B > 0 && h > 0 && w > 0, CSS_E_ARG, "css_rep_pass: non-positive size");
    CSS_CHECK_ARG(mode == CSS_SIM_COS || mode == CSS_SIM_SOFTMAX || (mode == 2 && css_rep_pass_use_tc()), CSS_E_ARG, "css_rep_pass: bad mode %d", mode);
    if (int e = css_check_dims(want_sim ? C : 1, D)) return e;
    CSS_CHECK_ARG(rep_dtype == CSS_DTYPE_F32 || rep_dtype == CSS_DTYPE_BF16, CSS_E_DTYPE, "css_rep_pass: rep dtype %d not supported",
                  rep_dtype);
    CSS_CHECK_ARG((long long)B * h * w < (1ll << 31) / CSS_CMAX, CSS_E_SIZE, "css_rep_pass: too many pixels");
    cudaStream_t st = (cudaStream_t)stream;
    const int hw = h * w, N = B * hw;
    int launches = 1;
    if (want_sim && rep_dtype == CSS_DTYPE_F32 && css_rep_pass_use_tc()) {       // tensor-core path (css_sim_tc.cu)
        if (int e = css_rep_pass_tc((const float*)rep, prototypes, proto_scratch, B, C, h, w, mode, temp, sim_out, (float*)rows, norms, st))
            return e;
        CSS_CHECK_LAUNCH("css_rep_pass", 2);
        return 0;
    }
    if (want_sim) {
        css_launch(proto_prep_kernel, dim3(CSS_CMAX), dim3(CSS_D), (size_t)(0), (cudaStream_t)(st), prototypes, proto_scratch, C);
        ++launches;
    }
#define REP_PASS_RUN(TYPE)                                                                                                  \
    do {                                                                                                                    \
        const TYPE* r = (const TYPE*)rep;                                                                                   \
        TYPE* rw = (TYPE*)rows;                                                                                             \
        if (want_sim && want_rows) dispatch_rep_pass<true, TYPE>((C + 3) / 4, r, proto_scratch, hw, N, C, mode, temp, sim_out, rw, norms, st); \
        else if (want_sim) dispatch_rep_pass<false, TYPE>((C + 3) / 4, r, proto_scratch, hw, N, C, mode, temp, sim_out, rw, norms, st);      \
        else launch_rep_pass<0, true, TYPE>(r, nullptr, hw, N, C, mode, temp, nullptr, rw, norms, st);                        \
    } while (0)
    if (rep_dtype == CSS_DTYPE_F32) REP_PASS_RUN(float);
    else REP_PASS_RUN(__nv_bfloat16);
#undef REP_PASS_RUN
    CSS_CHECK_LAUNCH("css_rep_pass", launches);
    return 0;
}

extern "C" int css_sim_map(const void* rep, int rep_dtype, const float* prototypes, float* proto_scratch, int B, int C,
                           int D, int h, int w, int mode, float temp, float* out, void* stream) {
    CSS_CHECK_ARG(rep && prototypes && proto_scratch && out, CSS_E_ARG, "css_sim_map: null pointer");
    return css_rep_pass(rep, rep_dtype, prototypes, proto_scratch, B, C, D, h, w, mode, temp, out, nullptr, nullptr, stream);
}

// ---------------------------------------------------------------------------------------------------------------
// css_rows_refresh: "are these pixel-major rows still the rows of THIS map?"  The student pass hands rows / norms to the loss
// on the `prob` tensor; the reference wraps the model in DistributedDataParallel(find_unused_parameters=True)
// (mix_label.py:77), whose output sink CLONES rep_all, so the loss sees equal content at another address.  Instead of
// re-reading the whole map (the rows-only pass), a sampled comparison decides on the device: RV_SAMPLES adjacent channels of
// EVERY pixel (the channel quad rotating with warp and round, so every channel plane is touched) are compared bit for bit with
// the carried rows, which raises meta[CSS_META_ROWS_STALE]; the rows-only pass that follows returns at once unless that word is
// set.  No host synchronisation, graph-capturable.  The word is cleared by css_select, which therefore runs first.
// ---------------------------------------------------------------------------------------------------------------
#define RV_SAMPLES 4
template <typename T>
__global__ void __launch_bounds__(256) rows_verify_kernel(const T* __restrict__ rep, const T* __restrict__ rows, int hw, int N,
                                                          int32_t* __restrict__ meta) {
    css_pdl_enter();
    // A warp owns 32 consecutive pixels and checks them in RV_SAMPLES rounds of 8 pixels x 4 ADJACENT channels (lane l: pixel
    // 8 i + l / 4, channel quad + l % 4): a round reads one 32-byte sector per pixel row and 4 planes x 32 bytes of the map -- 12
    // sectors for 32 samples, against 64 when every lane picked its own channel.  The quad rotates with warp and round.
    const int lane = threadIdx.x & 31, wbase = (blockIdx.x * 256 + threadIdx.x) & ~31;
    const unsigned wq = (unsigned)(wbase >> 5) * RV_SAMPLES;
    bool bad = false;
#pragma unroll
    for (int i = 0; i < RV_SAMPLES; ++i) {
        const int p = wbase + 8 * i + (lane >> 2);
        const int d = 4 * (int)(((wq + i) * 37u) & (unsigned)(CSS_D / 4 - 1)) + (lane & 3);
        if (p < N) {
            const int b = p / hw, s = p - b * hw;
            const T a = rep[((size_t)b * CSS_D + d) * hw + s];
            const T r = rows[(size_t)p * CSS_D + d];
            if constexpr (sizeof(T) == 4) bad |= __float_as_uint(a) != __float_as_uint(r);
            else bad |= __bfloat16_as_ushort(a) != __bfloat16_as_ushort(r);
        }
    }
    if (bad) meta[CSS_META_ROWS_STALE] = 1;
}

extern "C" int css_rows_refresh(const void* rep, int rep_dtype, void* rows, float* norms, int32_t* meta, int B, int D, int h, int w,
                                void* stream) {
    CSS_CHECK_ARG(rep && rows && norms && meta, CSS_E_ARG, "css_rows_refresh: null pointer");
    CSS_CHECK_ARG(B > 0 && h > 0 && w > 0, CSS_E_ARG, "css_rows_refresh: non-positive size");
    CSS_CHECK_ARG(D == CSS_D, CSS_E_DIM, "css_rows_refresh: D must be %d", CSS_D);
    CSS_CHECK_ARG(rep_dtype == CSS_DTYPE_F32 || rep_dtype == CSS_DTYPE_BF16, CSS_E_DTYPE, "css_rows_refresh: rep dtype %d not supported",
                  rep_dtype);
    CSS_CHECK_ARG((long long)B * h * w < (1ll << 31) / CSS_CMAX, CSS_E_SIZE, "css_rows_refresh: too many pixels");
    cudaStream_t st = (cudaStream_t)stream;
    const int hw = h * w, N = B * hw;
    const int32_t* guard = meta + CSS_META_ROWS_STALE;
    if (rep_dtype == CSS_DTYPE_F32) {
        css_launch(rows_verify_kernel<float>, dim3((N + 255) / 256), dim3(256), (size_t)(0), (cudaStream_t)(st), (const float*)rep, (const float*)rows, hw, N, meta);
        launch_rep_pass<0, true, float>((const float*)rep, nullptr, hw, N, 1, CSS_SIM_COS, 1.f, nullptr, (float*)rows, norms, st, guard);
    } else {
        css_launch(rows_verify_kernel<__nv_bfloat16>, dim3((N + 255) / 256), dim3(256), (size_t)(0), (cudaStream_t)(st), (const __nv_bfloat16*)rep, (const __nv_bfloat16*)rows, hw, N, meta);
        launch_rep_pass<0, true, __nv_bfloat16>((const __nv_bfloat16*)rep, nullptr, hw, N, 1, CSS_SIM_COS, 1.f, nullptr,
                                                (__nv_bfloat16*)rows, norms, st, guard);
    }
    CSS_CHECK_LAUNCH("css_rows_refresh", 2);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// K2: fused bilinear (align_corners=True) up-sampling + softmax + max (+ mix fusion) at crop resolution.
// A CTA owns an 8 x 32 tile of crop pixels of one image; the few low-resolution pixels that tile touches (for 81 -> 321:
// <= 5 x 11) are staged ONCE in shared memory, position-major ([pos][C rounded up to 4]) for both maps, so each of the
// 4 taps is one 128-bit shared load per four classes instead of 168 L1/L2 loads per thread, and the interpolation runs on
// packed fp32 pairs.  Arithmetic follows ATen's upsample_bilinear2d op for op (explicit
// _rn intrinsics: no FMA contraction):
//   src = ((in-1)/(out-1)) * dst ; i0 = int(src) ; i1 = i0 + (i0 < in-1) ; lam = src - i0
//   v = (1-ly)*((1-lx)*v00 + lx*v01) + ly*((1-lx)*v10 + lx*v11)
// Algorithmic bytes / crop pixel: 28 written (2 x f32 conf, 2 x i64 label, f32 fused) + the low-res maps read once.
// ---------------------------------------------------------------------------------------------------------------
#define K2_TH 8
#define K2_TW 32

struct Taps {
    int o00, o01, o10, o11;
    float hx, lx, hy, ly;
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// softmax + max over the C up-sampled values of one map; `src` is the staged tile ([C][cstride]) or the global map.
// CT > 0: compile-time class count (no predicated-off iterations); CT == 0: any C <= 32.
// tmode: 0 = no temperature, 1 = x * rtemp is exactly x / temp (temp a power of two), 2 = IEEE division by temp.
template <int CT>
__device__ __forceinline__ void upsample_softmax_max(const float* __restrict__ src, int C, int cstride, const Taps& t, int tmode,
                                                     float temp, float rtemp, float& conf, int& label) {
    constexpr int CN = CT > 0 ? CT : CSS_CMAX;
    float v[CN];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < CN; ++c) {
        if (CT > 0 || c < C) {
            const float* s = src + c * cstride;
            float top = __fadd_rn(__fmul_rn(t.hx, s[t.o00]), __fmul_rn(t.lx, s[t.o01]));
            float bot = __fadd_rn(__fmul_rn(t.hx, s[t.o10]), __fmul_rn(t.lx, s[t.o11]));
            float val = __fadd_rn(__fmul_rn(t.hy, top), __fmul_rn(t.ly, bot));
            if (tmode == 1) val = __fmul_rn(val, rtemp);
            else if (tmode == 2) val = __fdiv_rn(val, temp);
            v[c] = val;
            m = fmaxf(m, val);
        }
    }
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < CN; ++c) {
        if (CT > 0 || c < C) {
            v[c] = ex2_approx((v[c] - m) * 1.4426950408889634f);   // exp(x - m): exactly 1 for the arg-max, monotone elsewhere
            sum += v[c];
        }
    }
    // torch.max(softmax): the largest e_c/sum is e = 1 (the arg-max); first index on ties (ties after rounding included)
    const float emax = 1.f;
    int lab = 0;
#pragma unroll
    for (int c = CN - 1; c >= 0; --c)
        if ((CT > 0 || c < C) && v[c] == emax) lab = c;
    conf = __fdiv_rn(emax, sum);
    label = lab;
}

// Same result as upsample_softmax_max, for the staged tile: position-major [pos][CP] (CP = C rounded up to 4), so one
// 128-bit shared load per tap brings four classes and the interpolation runs on packed fp32 pairs (FMUL2 / FADD2: each
// half is the same IEEE round-to-nearest op as the scalar form, so labels stay bit-exact).  ~12 issue slots per class
// instead of ~32 (4 LDS + 4 address LEAs + 10 scalar flops + ...): the kernel is issue-bound, not HBM-bound.
template <int CT>
__device__ __forceinline__ void upsample_softmax_max_tile(const float* __restrict__ tile, int C, const Taps& t, int tmode, float temp,
                                                          float rtemp, float& conf, int& label) {
    constexpr int CP = CT > 0 ? ((CT + 3) & ~3) : CSS_CMAX;
    const float4* p00 = reinterpret_cast<const float4*>(tile + t.o00 * CP);
    const float4* p01 = reinterpret_cast<const float4*>(tile + t.o01 * CP);
    const float4* p10 = reinterpret_cast<const float4*>(tile + t.o10 * CP);
    const float4* p11 = reinterpret_cast<const float4*>(tile + t.o11 * CP);
    const float2 hx = make_float2(t.hx, t.hx), lx = make_float2(t.lx, t.lx);
    const float2 hy = make_float2(t.hy, t.hy), ly = make_float2(t.ly, t.ly);
    const float2 rt = make_float2(rtemp, rtemp);
    float2 v[CP / 2];
#pragma unroll
    for (int g = 0; g < CP / 4; ++g) {
        const float4 a = p00[g], b = p01[g], c = p10[g], d = p11[g];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const float2 a2 = hh ? make_float2(a.z, a.w) : make_float2(a.x, a.y);
            const float2 b2 = hh ? make_float2(b.z, b.w) : make_float2(b.x, b.y);
            const float2 c2 = hh ? make_float2(c.z, c.w) : make_float2(c.x, c.y);
            const float2 d2 = hh ? make_float2(d.z, d.w) : make_float2(d.x, d.y);
            const float2 top = __fadd2_rn(__fmul2_rn(hx, a2), __fmul2_rn(lx, b2));
            const float2 bot = __fadd2_rn(__fmul2_rn(hx, c2), __fmul2_rn(lx, d2));
            float2 val = __fadd2_rn(__fmul2_rn(hy, top), __fmul2_rn(ly, bot));
            if (tmode == 1) val = __fmul2_rn(val, rt);
            else if (tmode == 2) val = make_float2(__fdiv_rn(val.x, temp), __fdiv_rn(val.y, temp));
            v[2 * g + hh] = val;
        }
    }
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < CP / 2; ++i) {
        if (CT > 0 ? (2 * i < CT) : (2 * i < C)) m = fmaxf(m, v[i].x);
        if (CT > 0 ? (2 * i + 1 < CT) : (2 * i + 1 < C)) m = fmaxf(m, v[i].y);
    }
    const float2 nm = make_float2(-m, -m), l2e = make_float2(1.4426950408889634f, 1.4426950408889634f);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < CP / 2; ++i) {
        const float2 x = __fmul2_rn(__fadd2_rn(v[i], nm), l2e);   // (v - m) * log2(e), as the scalar form
        if (CT > 0 ? (2 * i < CT) : (2 * i < C)) sum += (v[i].x = ex2_approx(x.x));
        if (CT > 0 ? (2 * i + 1 < CT) : (2 * i + 1 < C)) sum += (v[i].y = ex2_approx(x.y));
    }
    int lab = 0;
#pragma unroll
    for (int cc = CP - 1; cc >= 0; --cc) {
        const float e = (cc & 1) ? v[cc >> 1].y : v[cc >> 1].x;
        if ((CT > 0 ? (cc < CT) : (cc < C)) && e == 1.f) lab = cc;
    }
    conf = __fdiv_rn(1.f, sum);
    label = lab;
}

template <bool STAGED, int CT>
__global__ void __launch_bounds__(K2_TH * K2_TW, STAGED ? 5 : 1) upsample_label_fuse_kernel(
    const float* __restrict__ sim, const float* __restrict__ logits, float temp, float rtemp, int tmode, int fuse_mode, int C, int h, int w, int H, int W,
    float ry, float rx, int tile_cap, float* __restrict__ conf_rep, int64_t* __restrict__ label_rep, float* __restrict__ conf_cls,
    int64_t* __restrict__ label_cls, float* __restrict__ fused) {
    css_pdl_enter();
    extern __shared__ __align__(16) float tile[];   // [2 maps][tile_cap positions][CP classes]  (STAGED only)
    const int b = blockIdx.z;
    const int Y0 = blockIdx.y * K2_TH, X0 = blockIdx.x * K2_TW;
    const int Y = Y0 + (threadIdx.x >> 5), X = X0 + (threadIdx.x & 31);
    const int hw = h * w;
    int ys0 = 0, xs0 = 0, in_tw = w, cstride = hw;
    if (STAGED) {
        const int Yl = min(Y0 + K2_TH - 1, H - 1), Xl = min(X0 + K2_TW - 1, W - 1);
        ys0 = (int)__fmul_rn(ry, (float)Y0);
        xs0 = (int)__fmul_rn(rx, (float)X0);
        const int ye = (int)__fmul_rn(ry, (float)Yl), xe = (int)__fmul_rn(rx, (float)Xl);
        const int in_th = min(ye + 1, h - 1) - ys0 + 1;
        in_tw = min(xe + 1, w - 1) - xs0 + 1;
        cstride = in_th * in_tw;                    // <= tile_cap by construction of the launch
        // transpose NCHW -> [pos][CP] through a 4-position x 8-class lane pattern: the global side reads 16-byte runs,
        // the shared side lands in 32 distinct banks for CP = 24 (stride 24 floats: rows 0,24,16,8 + 8 classes each)
        constexpr int CP = CT > 0 ? ((CT + 3) & ~3) : CSS_CMAX;
        // a warp walks 4-position x 8-class blocks; the only run-time division, r -> (yy, xx), is done in fp32 (exact: r < 2^12)
        constexpr int CB = (CP + 7) >> 3;
        const int RB = (cstride + 3) >> 2;
        const float inv_tw = __frcp_ru((float)in_tw);
        const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int blk = wid; blk < RB * CB; blk += K2_TH) {
            const int c_hi = blk % CB, r_hi = blk / CB;              // CB is a compile-time constant
            const int r = r_hi * 4 + (lane & 3), c = c_hi * 8 + (lane >> 2);
            if (r < cstride && c < CP) {
                int yy = (int)(((float)r + 0.5f) * inv_tw);
                yy -= (yy * in_tw > r);
                const int xx = r - yy * in_tw;
                const size_t g = ((size_t)b * C + c) * hw + (size_t)(ys0 + yy) * w + xs0 + xx;
                if (sim) tile[r * CP + c] = c < C ? __ldg(sim + g) : 0.f;
                if (logits) tile[(tile_cap + r) * CP + c] = c < C ? __ldg(logits + g) : 0.f;
            }
        }
        __syncthreads();
    }
    if (Y >= H || X >= W) return;
    Taps t;
    {
        const float ys = __fmul_rn(ry, (float)Y), xs = __fmul_rn(rx, (float)X);
        const int y0 = (int)ys, x0 = (int)xs;
        const int y1 = y0 + (y0 < h - 1), x1 = x0 + (x0 < w - 1);
        t.ly = __fsub_rn(ys, (float)y0);
        t.hy = __fsub_rn(1.f, t.ly);
        t.lx = __fsub_rn(xs, (float)x0);
        t.hx = __fsub_rn(1.f, t.lx);
        t.o00 = (y0 - ys0) * in_tw + (x0 - xs0);
        t.o01 = (y0 - ys0) * in_tw + (x1 - xs0);
        t.o10 = (y1 - ys0) * in_tw + (x0 - xs0);
        t.o11 = (y1 - ys0) * in_tw + (x1 - xs0);
    }
    const size_t o = ((size_t)b * H + Y) * W + X;
    int lr = -1, lc = -2;
    if (sim) {
        float cf;
        if (STAGED) upsample_softmax_max_tile<CT>(tile, C, t, tmode, temp, rtemp, cf, lr);
        else upsample_softmax_max<CT>(sim + (size_t)b * C * hw, C, hw, t, tmode, temp, rtemp, cf, lr);
        if (conf_rep) conf_rep[o] = cf;
        if (label_rep) label_rep[o] = lr;
    }
    if (logits) {
        float cf;
        if (STAGED) upsample_softmax_max_tile<CT>(tile + (size_t)tile_cap * (CT > 0 ? ((CT + 3) & ~3) : CSS_CMAX), C, t, 0, 1.f, 1.f, cf, lc);
        else upsample_softmax_max<CT>(logits + (size_t)b * C * hw, C, hw, t, 0, 1.f, 1.f, cf, lc);
        if (conf_cls) conf_cls[o] = cf;
        if (label_cls) label_cls[o] = lc;
    }
    if (fused && fuse_mode == CSS_FUSE_MIX) fused[o] = (lr == lc) ? (float)lc : 255.f;   // ddp_model.py:115-118
}

extern "C" int css_upsample_label_fuse(const float* sim, const float* logits, float temp, int fuse_mode, int B, int C, int h,
                                       int w, int H, int W, float* conf_rep, int64_t* label_rep, float* conf_cls,
                                       int64_t* label_cls, float* fused, void* stream) {
    CSS_CHECK_ARG(sim || logits, CSS_E_ARG, "css_upsample_label_fuse: both inputs null");
    CSS_CHECK_ARG(B > 0 && h > 0 && w > 0 && H > 0 && W > 0, CSS_E_ARG, "css_upsample_label_fuse: non-positive size");
    CSS_CHECK_ARG(C >= 1 && C <= CSS_CMAX, CSS_E_DIM, "css_upsample_label_fuse: C must be in [1,%d]", CSS_CMAX);
    CSS_CHECK_ARG(sim || (!conf_rep && !label_rep), CSS_E_ARG, "css_upsample_label_fuse: rep outputs without sim");
    CSS_CHECK_ARG(logits || (!conf_cls && !label_cls), CSS_E_ARG, "css_upsample_label_fuse: cls outputs without logits");
    CSS_CHECK_ARG(fuse_mode == CSS_FUSE_NONE || (fuse_mode == CSS_FUSE_MIX && sim && logits && fused), CSS_E_ARG,
                  "css_upsample_label_fuse: mix fusion needs sim, logits and fused");
    CSS_CHECK_ARG(B <= 65535 && (H + K2_TH - 1) / K2_TH <= 65535 && (long long)B * C * h * w < (1ll << 31), CSS_E_SIZE,
                  "css_upsample_label_fuse: B, H or the map too large");
    // area_pixel_compute_scale(align_corners=True): (in-1)/(out-1) in fp32, 0 when out == 1
    const float ry = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f;
    const float rx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
    // upper bound of the low-res footprint of one 8 x 32 output tile (+1 of slack for fp32 rounding of src indices)
    const int in_th = (int)fminf((float)h, ceilf(ry * (K2_TH - 1)) + 3.f);
    const int in_tw = (int)fminf((float)w, ceilf(rx * (K2_TW - 1)) + 3.f);
    const int tile_cap = in_th * in_tw;
    const int cp = (C == 21 || C == 19) ? ((C + 3) & ~3) : CSS_CMAX;       // class pitch of the staged tile (see the kernel)
    const size_t smem = (size_t)2 * cp * tile_cap * sizeof(float);
    dim3 grid((W + K2_TW - 1) / K2_TW, (H + K2_TH - 1) / K2_TH, B);
    cudaStream_t st = (cudaStream_t)stream;
    // x / temp == x * (1/temp) exactly when temp is a power of two (the shipped 0.5 / 0.25); otherwise divide like the reference
    int texp;
    const int tmode = (frexpf(temp, &texp) == 0.5f) ? 1 : 2;
    const float rtemp = 1.f / temp;
#define K2_LAUNCH(STG, CTV, SM)                                                                                                     \
    css_launch(upsample_label_fuse_kernel<STG, CTV>, dim3(grid), dim3(K2_TH * K2_TW), (size_t)(SM), (cudaStream_t)(st), sim, logits, temp, rtemp, tmode, fuse_mode, C, h, w, H, W, ry, \
                                                                         rx, tile_cap, conf_rep, label_rep, conf_cls, label_cls, fused)
    if (smem <= 96 * 1024) {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(upsample_label_fuse_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(upsample_label_fuse_kernel<true, 19>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(upsample_label_fuse_kernel<true, 21>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
            if (e != cudaSuccess) { css_set_error("css_upsample_label_fuse: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
        }
        if (C == 21) K2_LAUNCH(true, 21, smem);          // VOC
        else if (C == 19) K2_LAUNCH(true, 19, smem);     // CityScapes
        else K2_LAUNCH(true, 0, smem);
    } else {   // strong down-sampling: the footprint does not fit, read the taps from global memory
        K2_LAUNCH(false, 0, 0);
    }
#undef K2_LAUNCH
    CSS_CHECK_LAUNCH("css_upsample_label_fuse", 1);
    return 0;
}
