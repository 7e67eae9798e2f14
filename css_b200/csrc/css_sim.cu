// Stage 1 / 1b / 1' / 2: similarity maps against the class prototypes, fused up-sample + softmax + max + fusion.
// Reference: generalframeworks/networks/ddp_model.py:104-118, 147-154 (Model_mix); :189-199, 230-237 (Model_cross);
// :36-37 (Model_ori_pseudo).
#include "css_common.cuh"

// ---------------------------------------------------------------------------------------------------------------
// prototype preparation: F.normalize(prototypes, dim=-1) (eps 1e-12, ddp_model.py:107), transposed to [D][32]
// (class-minor, zero padded) so the streaming kernel reads 4 classes per 128-bit shared-memory load.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) proto_prep_kernel(const float* __restrict__ protos, float* __restrict__ scratch, int C) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __shared__ float nrm[CSS_CMAX];
    for (int c = warp; c < CSS_CMAX; c += 8) {
        float s = 0.f;
        if (c < C)
            for (int d = lane; d < CSS_D; d += 32) {
                float v = protos[c * CSS_D + d];
                s = fmaf(v, v, s);
            }
        s = warp_sum(s);
        if (lane == 0) nrm[c] = (c < C) ? fmaxf(sqrtf(s), 1e-12f) : 1.f;
    }
    __syncthreads();
    const int d = threadIdx.x;
    for (int c = 0; c < CSS_CMAX; ++c)
        scratch[d * CSS_CMAX + c] = (c < C) ? __fdiv_rn(protos[c * CSS_D + d], nrm[c]) : 0.f;
}

// ---------------------------------------------------------------------------------------------------------------
// K1: one thread per pixel streams its 256 channels straight from the NCHW map (a warp reads 128 contiguous bytes
// per channel), keeps 4*NG class dots + the squared norm in registers, prototypes broadcast from shared memory.
//   cos_c = (x . p_hat_c) / max(||x||, 1e-12)                                   (ddp_model.py:105-109)
//   mode CSS_SIM_SOFTMAX: softmax_c(cos_c / temp)                               (ddp_model.py:154)
// Algorithmic bytes / pixel: D*4 read + C*4 written.
// ---------------------------------------------------------------------------------------------------------------
template <int NG>
__global__ void __launch_bounds__(128) sim_map_kernel(const float* __restrict__ rep, const float* __restrict__ scratch,
                                                      int hw, int N, int C, int mode, float temp, float* __restrict__ out) {
    __shared__ float4 sp[CSS_D * NG];
    for (int i = threadIdx.x; i < CSS_D * NG; i += 128) {
        int d = i / NG, g = i - d * NG;
        sp[i] = reinterpret_cast<const float4*>(scratch)[d * (CSS_CMAX / 4) + g];
    }
    __syncthreads();
    const int p = blockIdx.x * 128 + threadIdx.x;
    if (p >= N) return;
    const int b = p / hw, s = p - b * hw;
    const float* x = rep + (size_t)b * CSS_D * hw + s;

    float acc[4 * NG];
#pragma unroll
    for (int i = 0; i < 4 * NG; ++i) acc[i] = 0.f;
    float n2 = 0.f;
    constexpr int U = 8;
    for (int d0 = 0; d0 < CSS_D; d0 += U) {
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ldg_stream(x + (size_t)(d0 + u) * hw);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            n2 = fmaf(v[u], v[u], n2);
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                float4 q = sp[(d0 + u) * NG + g];
                acc[4 * g + 0] = fmaf(v[u], q.x, acc[4 * g + 0]);
                acc[4 * g + 1] = fmaf(v[u], q.y, acc[4 * g + 1]);
                acc[4 * g + 2] = fmaf(v[u], q.z, acc[4 * g + 2]);
                acc[4 * g + 3] = fmaf(v[u], q.w, acc[4 * g + 3]);
            }
        }
    }
    const float nrm = fmaxf(sqrtf(n2), 1e-12f);
    float* o = out + (size_t)b * C * hw + s;
    if (mode == CSS_SIM_COS) {
#pragma unroll
        for (int c = 0; c < 4 * NG; ++c)
            if (c < C) o[(size_t)c * hw] = __fdiv_rn(acc[c], nrm);
    } else {
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4 * NG; ++c) {
            acc[c] = __fdiv_rn(__fdiv_rn(acc[c], nrm), temp);
            if (c < C) m = fmaxf(m, acc[c]);
        }
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < 4 * NG; ++c) {
            acc[c] = (c < C) ? expf(acc[c] - m) : 0.f;
            sum += acc[c];
        }
#pragma unroll
        for (int c = 0; c < 4 * NG; ++c)
            if (c < C) o[(size_t)c * hw] = __fdiv_rn(acc[c], sum);
    }
}

template <int NG>
static void launch_sim(const float* rep, const float* scratch, int hw, int N, int C, int mode, float temp, float* out,
                       cudaStream_t st) {
    sim_map_kernel<NG><<<(N + 127) / 128, 128, 0, st>>>(rep, scratch, hw, N, C, mode, temp, out);
}

extern "C" int css_sim_map(const void* rep, int rep_dtype, const float* prototypes, float* proto_scratch, int B, int C,
                           int D, int h, int w, int mode, float temp, float* out, void* stream) {
    CSS_CHECK_ARG(rep && prototypes && proto_scratch && out, CSS_E_ARG, "css_sim_map: null pointer");
    CSS_CHECK_ARG(B > 0 && h > 0 && w > 0, CSS_E_ARG, "css_sim_map: non-positive size");
    CSS_CHECK_ARG(mode == CSS_SIM_COS || mode == CSS_SIM_SOFTMAX, CSS_E_ARG, "css_sim_map: bad mode %d", mode);
    if (int e = css_check_dims(C, D)) return e;
    CSS_CHECK_ARG(rep_dtype == CSS_DTYPE_F32, CSS_E_DTYPE, "css_sim_map: rep dtype %d not supported", rep_dtype);
    CSS_CHECK_ARG((long long)B * h * w < (1ll << 31) / CSS_CMAX, CSS_E_SIZE, "css_sim_map: too many pixels");
    cudaStream_t st = (cudaStream_t)stream;
    const int hw = h * w, N = B * hw;
    proto_prep_kernel<<<1, 256, 0, st>>>(prototypes, proto_scratch, C);
    const float* r = (const float*)rep;
    switch ((C + 3) / 4) {
        case 1: launch_sim<1>(r, proto_scratch, hw, N, C, mode, temp, out, st); break;
        case 2: launch_sim<2>(r, proto_scratch, hw, N, C, mode, temp, out, st); break;
        case 3: launch_sim<3>(r, proto_scratch, hw, N, C, mode, temp, out, st); break;
        case 4: launch_sim<4>(r, proto_scratch, hw, N, C, mode, temp, out, st); break;
        case 5: launch_sim<5>(r, proto_scratch, hw, N, C, mode, temp, out, st); break;
        case 6: launch_sim<6>(r, proto_scratch, hw, N, C, mode, temp, out, st); break;
        case 7: launch_sim<7>(r, proto_scratch, hw, N, C, mode, temp, out, st); break;
        default: launch_sim<8>(r, proto_scratch, hw, N, C, mode, temp, out, st); break;
    }
    CSS_CHECK_LAUNCH("css_sim_map", 2);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// K2: one thread per crop pixel.  4-tap align_corners=True bilinear of the C similarities and the C logits read from
// the rep-resolution maps (L2 resident), softmax(/temp) + max with first-index tie breaking, mix fusion.  Arithmetic
// follows ATen's upsample_bilinear2d op for op (explicit _rn intrinsics: no FMA contraction):
//   src = ((in-1)/(out-1)) * dst ; i0 = int(src) ; i1 = i0 + (i0 < in-1) ; lam = src - i0
//   v = (1-ly)*((1-lx)*v00 + lx*v01) + ly*((1-lx)*v10 + lx*v11)
// Algorithmic bytes / crop pixel: 28 written (2 x f32 conf, 2 x i64 label, f32 fused); the taps come from L2.
// ---------------------------------------------------------------------------------------------------------------
struct Taps {
    int o00, o01, o10, o11;
    float hx, lx, hy, ly;
};

template <bool USE_TEMP>
__device__ __forceinline__ void upsample_softmax_max(const float* __restrict__ src, int C, int hw, const Taps& t, float temp,
                                                     float& conf, int& label) {
    float v[CSS_CMAX];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < CSS_CMAX; ++c) {
        if (c < C) {
            const float* s = src + (size_t)c * hw;
            float top = __fadd_rn(__fmul_rn(t.hx, __ldg(s + t.o00)), __fmul_rn(t.lx, __ldg(s + t.o01)));
            float bot = __fadd_rn(__fmul_rn(t.hx, __ldg(s + t.o10)), __fmul_rn(t.lx, __ldg(s + t.o11)));
            float val = __fadd_rn(__fmul_rn(t.hy, top), __fmul_rn(t.ly, bot));
            if (USE_TEMP) val = __fdiv_rn(val, temp);
            v[c] = val;
            m = fmaxf(m, val);
        }
    }
    float sum = 0.f;
    float emax = 0.f;
#pragma unroll
    for (int c = 0; c < CSS_CMAX; ++c) {
        if (c < C) {
            v[c] = expf(v[c] - m);
            sum += v[c];
            emax = fmaxf(emax, v[c]);
        }
    }
    // torch.max(softmax): the largest e_c/sum, first index on ties (ties after rounding included)
    int lab = 0;
#pragma unroll
    for (int c = CSS_CMAX - 1; c >= 0; --c)
        if (c < C && v[c] == emax) lab = c;
    conf = __fdiv_rn(emax, sum);
    label = lab;
}

__global__ void __launch_bounds__(128) upsample_label_fuse_kernel(const float* __restrict__ sim, const float* __restrict__ logits,
                                                                  float temp, int fuse_mode, int C, int h, int w, int H, int W,
                                                                  float ry, float rx, float* __restrict__ conf_rep,
                                                                  int64_t* __restrict__ label_rep, float* __restrict__ conf_cls,
                                                                  int64_t* __restrict__ label_cls, float* __restrict__ fused) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (idx >= H * W) return;
    const int Y = idx / W, X = idx - Y * W;
    Taps t;
    {
        const float ys = __fmul_rn(ry, (float)Y), xs = __fmul_rn(rx, (float)X);
        const int y0 = (int)ys, x0 = (int)xs;
        const int y1 = y0 + (y0 < h - 1), x1 = x0 + (x0 < w - 1);
        t.ly = __fsub_rn(ys, (float)y0);
        t.hy = __fsub_rn(1.f, t.ly);
        t.lx = __fsub_rn(xs, (float)x0);
        t.hx = __fsub_rn(1.f, t.lx);
        t.o00 = y0 * w + x0;
        t.o01 = y0 * w + x1;
        t.o10 = y1 * w + x0;
        t.o11 = y1 * w + x1;
    }
    const int hw = h * w;
    const size_t o = ((size_t)b * H + Y) * W + X;
    int lr = -1, lc = -2;
    if (sim) {
        float c;
        upsample_softmax_max<true>(sim + (size_t)b * C * hw, C, hw, t, temp, c, lr);
        if (conf_rep) conf_rep[o] = c;
        if (label_rep) label_rep[o] = lr;
    }
    if (logits) {
        float c;
        upsample_softmax_max<false>(logits + (size_t)b * C * hw, C, hw, t, 1.f, c, lc);
        if (conf_cls) conf_cls[o] = c;
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
    CSS_CHECK_ARG(B <= 65535 && (long long)H * W < (1ll << 31), CSS_E_SIZE, "css_upsample_label_fuse: B or H*W too large");
    // area_pixel_compute_scale(align_corners=True): (in-1)/(out-1) in fp32, 0 when out == 1
    const float ry = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f;
    const float rx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
    dim3 grid((H * W + 127) / 128, B);
    upsample_label_fuse_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(sim, logits, temp, fuse_mode, C, h, w, H, W, ry, rx,
                                                                       conf_rep, label_rep, conf_cls, label_cls, fused);
    CSS_CHECK_LAUNCH("css_upsample_label_fuse", 1);
    return 0;
}
