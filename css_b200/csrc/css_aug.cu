// the label / confidence maps' trip through the augmentation, kept on the GPU.
// Replaces, for the maps only, tensor_to_pil_* + transform_* + the torch.cat(...).to(device) of batch_transform_*
// (dataset_helpers/VOC.py:64-352) and the per-image loop of generate_cut_gather_* (VOC.py:354-477).  The image keeps
// the reference's PIL path; the host draws the geometry there and hands 5 ints per image to these kernels.
#include "css_common.cuh"

#define AUG_GEO 5            // resized_h, resized_w, top, left, flip

// Pillow's NEAREST resize (affine scale path) walks `pos = 0.5 * a; pos += a` in double, a = n_in / n_out, and truncates.
// The running sum is what makes it differ from (x + 0.5) * a in the last bit, so it is replayed sequentially: one thread
// per (image, axis), <= a few thousand dependent DADDs (~10 us, off the critical path of anything else).
__global__ void aug_index_kernel(const int* __restrict__ geo, int H, int W, int max_r, int* __restrict__ ymap,
                                 int* __restrict__ xmap) {
    css_pdl_enter();
    const int b = blockIdx.x, axis = blockIdx.y;
    if (threadIdx.x != 0) return;
    const int n_out = geo[b * AUG_GEO + axis], n_in = axis ? W : H;
    int* map = (axis ? xmap : ymap) + (size_t)b * max_r;
    const double a = (double)n_in / (double)n_out;
    double pos = 0.0 + a * 0.5;
    for (int x = 0; x < n_out && x < max_r; ++x) {
        map[x] = min((int)pos, n_in - 1);
        pos += a;
    }
}

template <typename LT>
__device__ __forceinline__ int label_byte(LT v) {
    // (label.float() / 255) -> mul(255).byte(): identity on 0..255 (checked exhaustively in the oracle), -1 wraps to 255
    return ((int)v) & 255;
}

template <typename LT>
__global__ void __launch_bounds__(256) aug_maps_kernel(const LT* __restrict__ la, const LT* __restrict__ lb,
                                                       const float* __restrict__ ca, const float* __restrict__ cb,
                                                       const int* __restrict__ geo, const int* __restrict__ ymap,
                                                       const int* __restrict__ xmap, int H, int W, int max_r, int ch, int cw,
                                                       int64_t* __restrict__ ola, int64_t* __restrict__ olb,
                                                       float* __restrict__ oca, float* __restrict__ ocb) {
    css_pdl_enter();
    const int b = blockIdx.z, y = blockIdx.y, x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= cw) return;
    const int* g = geo + b * AUG_GEO;
    const int rh = g[0], rw = g[1], top = g[2], left = g[3], flip = g[4];
    const int ry = top + y, rx = left + (flip ? cw - 1 - x : x);
    const bool inside = ry < rh && rx < rw;                      // else: bottom / right constant padding
    size_t src = 0;
    if (inside) src = ((size_t)b * H + ymap[(size_t)b * max_r + ry]) * W + xmap[(size_t)b * max_r + rx];
    const size_t dst = ((size_t)b * ch + y) * cw + x;
    if (la) {
        const int v = inside ? label_byte(la[src]) : 255;
        ola[dst] = v == 255 ? -1 : v;
    }
    if (lb) {
        const int v = inside ? label_byte(lb[src]) : 255;
        olb[dst] = v == 255 ? -1 : v;
    }
    // to_pil_image: mul(255).byte() truncates; to_tensor: byte / 255 in fp32
    if (ca) oca[dst] = inside ? __fdiv_rn((float)(((int)__fmul_rn(ca[src], 255.f)) & 255), 255.f) : 0.f;
    if (cb) ocb[dst] = inside ? __fdiv_rn((float)(((int)__fmul_rn(cb[src], 255.f)) & 255), 255.f) : 0.f;
}

extern "C" int css_aug_index(const int32_t* geometry, int B, int H, int W, int max_r, int32_t* ymap, int32_t* xmap,
                             void* stream) {
    CSS_CHECK_ARG(geometry && ymap && xmap, CSS_E_ARG, "css_aug_index: null pointer");
    CSS_CHECK_ARG(B > 0 && B <= 65535 && H > 0 && W > 0 && max_r > 0, CSS_E_ARG, "css_aug_index: non-positive size");
    css_launch(aug_index_kernel, dim3(dim3(B, 2)), dim3(32), (size_t)(0), (cudaStream_t)((cudaStream_t)stream), geometry, H, W, max_r, ymap, xmap);
    CSS_CHECK_LAUNCH("css_aug_index", 1);
    return 0;
}

extern "C" int css_aug_maps(const void* label_a, const void* label_b, int label_dtype, const float* conf_a, const float* conf_b,
                            const int32_t* geometry, const int32_t* ymap, const int32_t* xmap, int B, int H, int W, int max_r,
                            int ch, int cw, int64_t* out_label_a, int64_t* out_label_b, float* out_conf_a, float* out_conf_b,
                            void* stream) {
    CSS_CHECK_ARG(geometry && ymap && xmap, CSS_E_ARG, "css_aug_maps: null pointer");
    CSS_CHECK_ARG(label_a || label_b || conf_a || conf_b, CSS_E_ARG, "css_aug_maps: no map given");
    CSS_CHECK_ARG((!label_a || out_label_a) && (!label_b || out_label_b) && (!conf_a || out_conf_a) && (!conf_b || out_conf_b),
                  CSS_E_ARG, "css_aug_maps: a map without its output");
    CSS_CHECK_ARG(B > 0 && B <= 65535 && H > 0 && W > 0 && max_r > 0 && ch > 0 && ch <= 65535 && cw > 0, CSS_E_ARG,
                  "css_aug_maps: bad size");
    CSS_CHECK_ARG(label_dtype == CSS_LABEL_F32 || label_dtype == CSS_LABEL_I64, CSS_E_DTYPE, "css_aug_maps: label dtype %d not supported",
                  label_dtype);
    const dim3 grid((cw + 255) / 256, ch, B);
    cudaStream_t st = (cudaStream_t)stream;
    if (label_dtype == CSS_LABEL_F32)
        css_launch(aug_maps_kernel<float>, dim3(grid), dim3(256), (size_t)(0), (cudaStream_t)(st), (const float*)label_a, (const float*)label_b, conf_a, conf_b, geometry, ymap, xmap,
                                                     H, W, max_r, ch, cw, out_label_a, out_label_b, out_conf_a, out_conf_b);
    else
        css_launch(aug_maps_kernel<int64_t>, dim3(grid), dim3(256), (size_t)(0), (cudaStream_t)(st), (const int64_t*)label_a, (const int64_t*)label_b, conf_a, conf_b, geometry, ymap,
                                                       xmap, H, W, max_r, ch, cw, out_label_a, out_label_b, out_conf_a, out_conf_b);
    CSS_CHECK_LAUNCH("css_aug_maps", 1);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// CutOut / CutMix / ClassMix in one launch.  keep = pixel outside the drawn box (cutout, cutmix) or own label_a in the
// drawn class set (classmix); out = keep ? own[i] : partner[(i + 1) % B]  (cutout: image / conf 0, label_a -1).
// The reference's `x * mask + y * (1 - mask)` with a 0/1 mask selects exactly like this.
// ---------------------------------------------------------------------------------------------------------------
struct CutMaps {
    const float* image;
    const int64_t *la, *lb;
    const float *ca, *cb;
};

__global__ void __launch_bounds__(256) cut_mix_kernel(CutMaps own, CutMaps par, const int* __restrict__ boxes,
                                                      const unsigned long long* __restrict__ class_sets, int mode, int B, int CH,
                                                      int H, int W, float* __restrict__ o_img, int64_t* __restrict__ o_la,
                                                      int64_t* __restrict__ o_lb, float* __restrict__ o_ca, float* __restrict__ o_cb) {
    css_pdl_enter();
    const int i = blockIdx.z, y = blockIdx.y, x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    const int j = (i + 1) % B;
    const size_t hw = (size_t)H * W, px = (size_t)y * W + x;
    const size_t mo = i * hw + px, mp = j * hw + px;
    bool keep;
    if (mode == CSS_CUT_CLASSMIX) {
        const long long v = own.la[mo];
        keep = v >= -1 && v <= 62 && ((class_sets[i] >> (int)(v + 1)) & 1ull);
    } else {
        const int* bx = boxes + 4 * i;
        keep = !(y >= bx[0] && y < bx[1] && x >= bx[2] && x < bx[3]);
    }
    const bool cutout = mode == CSS_CUT_CUTOUT;
    for (int c = 0; c < CH; ++c) {
        const size_t o = ((size_t)i * CH + c) * hw + px;
        o_img[o] = keep ? own.image[o] : (cutout ? 0.f : par.image[((size_t)j * CH + c) * hw + px]);
    }
    o_la[mo] = keep ? own.la[mo] : (cutout ? -1 : par.la[mp]);
    if (own.lb) o_lb[mo] = (keep || cutout) ? own.lb[mo] : par.lb[mp];
    o_ca[mo] = keep ? own.ca[mo] : (cutout ? 0.f : par.ca[mp]);
    if (own.cb) o_cb[mo] = keep ? own.cb[mo] : (cutout ? 0.f : par.cb[mp]);
}

extern "C" int css_cut_mix(const float* image, const int64_t* label_a, const int64_t* label_b, const float* conf_a,
                           const float* conf_b, const float* p_image, const int64_t* p_label_a, const int64_t* p_label_b,
                           const float* p_conf_a, const float* p_conf_b, const int32_t* boxes, const uint64_t* class_sets, int mode,
                           int B, int CH, int H, int W, float* out_image, int64_t* out_label_a, int64_t* out_label_b,
                           float* out_conf_a, float* out_conf_b, void* stream) {
    CSS_CHECK_ARG(image && label_a && conf_a && out_image && out_label_a && out_conf_a, CSS_E_ARG, "css_cut_mix: null pointer");
    CSS_CHECK_ARG(mode == CSS_CUT_CUTOUT || mode == CSS_CUT_CUTMIX || mode == CSS_CUT_CLASSMIX, CSS_E_ARG, "css_cut_mix: bad mode %d", mode);
    CSS_CHECK_ARG(mode == CSS_CUT_CLASSMIX ? class_sets != nullptr : boxes != nullptr, CSS_E_ARG,
                  "css_cut_mix: the mode's mask description (boxes / class_sets) is missing");
    CSS_CHECK_ARG(mode == CSS_CUT_CUTOUT || (p_image && p_label_a && p_conf_a && (!label_b || p_label_b) && (!conf_b || p_conf_b)),
                  CSS_E_ARG, "css_cut_mix: partner maps missing");
    CSS_CHECK_ARG((!label_b || out_label_b) && (!conf_b || out_conf_b), CSS_E_ARG, "css_cut_mix: a map without its output");
    CSS_CHECK_ARG(B > 0 && B <= 65535 && CH > 0 && H > 0 && H <= 65535 && W > 0, CSS_E_ARG, "css_cut_mix: bad size");
    CutMaps own{image, label_a, label_b, conf_a, conf_b}, par{p_image, p_label_a, p_label_b, p_conf_a, p_conf_b};
    css_launch(cut_mix_kernel, dim3(dim3((W + 255) / 256, H, B)), dim3(256), (size_t)(0), (cudaStream_t)((cudaStream_t)stream), own, par, boxes, (const unsigned long long*)class_sets,
                                                                                   mode, B, CH, H, W, out_image, out_label_a, out_label_b,
                                                                                   out_conf_a, out_conf_b);
    CSS_CHECK_LAUNCH("css_cut_mix", 1);
    return 0;
}
