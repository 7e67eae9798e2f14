/* TEST INFRASTRUCTURE ONLY -- a second, independent restatement (plain C, scalar loops) of the byte / integer / index work of
 * the path, used by tests/test_oracle_c.py to cross-check oracle/css_oracle.py and the golden bundles.  Nothing in css_b200/
 * links or calls it.  Reference lines are relative to the root of WangChangqi98/CSS.
 *
 *   gcc -O2 -shared -fPIC -o oracle/_build/libcss_oracle_int.so oracle/css_oracle_int.c -lm      (oracle/build_c.py)
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* Image.resize((.., n_out), NEAREST) along one axis: Pillow's affine scale path walks pos = 0.5*a, pos += a in double
 * (a = n_in / n_out) and truncates.  Third-party code (Pillow, not vendored); pinned against Pillow itself in the tests. */
void orc_pil_nearest_table(int n_in, int n_out, int64_t* out) {
    const double a = (double)n_in / (double)n_out;
    double pos = 0.0 + a * 0.5;
    for (int x = 0; x < n_out; ++x) {
        int64_t v = (int64_t)pos;
        out[x] = v < n_in - 1 ? v : n_in - 1;
        pos += a;
    }
}

/* tensor_to_pil_* + transform_* + to_tensor for ONE label map (dataset_helpers/VOC.py:126-196, :284-291):
 * label -> byte (value & 255: -1 and 255 both mean ignore) -> NEAREST resize to (rh, rw) -> bottom/right pad with 255 -> crop at
 * (top, left) -> optional hflip -> int64 with 255 -> -1. */
void orc_aug_label(const int64_t* label, int H, int W, int rh, int rw, int top, int left, int flip, int ch, int cw, int64_t* out) {
    int64_t ymap[8192], xmap[8192];
    orc_pil_nearest_table(H, rh, ymap);
    orc_pil_nearest_table(W, rw, xmap);
    for (int y = 0; y < ch; ++y)
        for (int x = 0; x < cw; ++x) {
            const int xs = flip ? cw - 1 - x : x;
            const int ry = top + y, rx = left + xs;
            int b = 255;
            if (ry < rh && rx < rw) b = (int)(label[ymap[ry] * W + xmap[rx]] & 255);
            out[(int64_t)y * cw + x] = b == 255 ? -1 : b;
        }
}

/* same trip for a confidence map: to_pil_image truncates conf * 255 to a byte, to_tensor divides by 255; padding is 0 */
void orc_aug_conf(const float* conf, int H, int W, int rh, int rw, int top, int left, int flip, int ch, int cw, float* out) {
    int64_t ymap[8192], xmap[8192];
    orc_pil_nearest_table(H, rh, ymap);
    orc_pil_nearest_table(W, rw, xmap);
    for (int y = 0; y < ch; ++y)
        for (int x = 0; x < cw; ++x) {
            const int xs = flip ? cw - 1 - x : x;
            const int ry = top + y, rx = left + xs;
            float v = 0.f;
            if (ry < rh && rx < rw) {
                const float scaled = conf[ymap[ry] * W + xmap[rx]] * 255.0f;
                const int byte = (int)((int64_t)scaled & 255);
                v = (float)byte / 255.0f;
            }
            out[(int64_t)y * cw + x] = v;
        }
}

/* ddp_model.py:115-118: keep the logit-space label where both spaces agree, else 255 (stored as float) */
void orc_mix_fuse(const int64_t* label_cls, const int64_t* label_rep, int64_t n, float* out) {
    for (int64_t i = 0; i < n; ++i) out[i] = label_cls[i] == label_rep[i] ? (float)label_cls[i] : 255.0f;
}

/* loss.py:80,93-99,111-113: valid_c = label[:,c] * mask != 0, hard_c = valid_c and prob[:,c] < strong; per-class lists of
 * pixel ids (b*h*w + y*w + x) in row-major order.  Lists have room for N entries per class. */
void orc_select(const float* label, const float* mask, const float* prob, float strong, int B2, int C, int h, int w, int32_t* valid_list,
                int32_t* hard_list, int32_t* n_valid, int32_t* n_hard) {
    const int hw = h * w;
    const int64_t N = (int64_t)B2 * hw;
    for (int c = 0; c < C; ++c) {
        int nv = 0, nh = 0;
        for (int b = 0; b < B2; ++b)
            for (int s = 0; s < hw; ++s) {
                const int64_t o = ((int64_t)b * C + c) * hw + s;
                const float v = label[o] * mask[(int64_t)b * hw + s];
                if (v != 0.0f) {
                    valid_list[c * N + nv++] = (int32_t)(b * hw + s);
                    if (prob[o] < strong) hard_list[c * N + nh++] = (int32_t)(b * hw + s);
                }
            }
        n_valid[c] = nv;
        n_hard[c] = nh;
    }
}

/* F.interpolate(mode='nearest') source index: min(floor(dst * (in / out)), in - 1) with the scale in fp32 (ATen) */
static int nearest_src(int dst, int in, int out) {
    const float scale = (float)in / (float)out;
    const int s = (int)floorf((float)dst * scale);
    return s < in - 1 ? s : in - 1;
}

/* mix_label.py:175-183 (mode 1), cross_label.py:178-185 / ori_pseudo.py:171-178 (mode 0); utils.py:116-136.
 * label_all [2B,C,h,w], mask_all [2B,1,h,w], both f32, from label_l / label_u [B,H,W] int64 and conf_u [B,H,W]. */
void orc_threshold_glue(const int64_t* label_l, const int64_t* label_u, const float* conf_u, float weak, int mode, int B, int C, int H,
                        int W, int h, int w, float* label_all, float* mask_all) {
    memset(label_all, 0, sizeof(float) * (size_t)2 * B * C * h * w);
    for (int bb = 0; bb < 2 * B; ++bb)
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                const int Y = nearest_src(y, H, h), X = nearest_src(x, W, w);
                const int64_t src = ((int64_t)(bb % B) * H + Y) * W + X;
                int cls;
                float m;
                if (bb < B) {
                    const int64_t l = label_l[src];
                    m = l >= 0 ? 1.0f : 0.0f;
                    cls = l > 0 ? (int)l : 0;                       /* label_onehot: relu, so -1 lands in class 0 */
                } else {
                    const int64_t l = label_u[src];
                    m = conf_u[src] >= weak ? 1.0f : 0.0f;
                    cls = mode == 1 ? (int)l : (l > 0 ? (int)l : 0); /* label_onehot_2: -1 -> dropped channel, i.e. all zero */
                }
                mask_all[((int64_t)bb * h + y) * w + x] = m;
                if (cls >= 0 && cls < C) label_all[(((int64_t)bb * C + cls) * h + y) * w + x] = 1.0f;
            }
}

/* generate_cut_gather_* for one rank's slice (VOC.py:354-477), box modes: keep = outside boxes[i] = (y0, y1, x0, x1);
 * out = keep ? own[i] : partner[(i+1) % B]; cutout (mode 0) writes 0 / -1 instead of the partner (label_b untouched). */
void orc_cut_mix_boxes(const float* image, const int64_t* label, const float* conf, const float* p_image, const int64_t* p_label,
                       const float* p_conf, const int32_t* boxes, int cutout, int B, int CH, int H, int W, float* o_image,
                       int64_t* o_label, float* o_conf) {
    const int64_t hw = (int64_t)H * W;
    for (int i = 0; i < B; ++i) {
        const int j = (i + 1) % B;
        const int32_t* bx = boxes + 4 * i;
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                const int keep = !(y >= bx[0] && y < bx[1] && x >= bx[2] && x < bx[3]);
                const int64_t px = (int64_t)y * W + x;
                for (int c = 0; c < CH; ++c)
                    o_image[((int64_t)i * CH + c) * hw + px] =
                        keep ? image[((int64_t)i * CH + c) * hw + px] : (cutout ? 0.0f : p_image[((int64_t)j * CH + c) * hw + px]);
                o_label[i * hw + px] = keep ? label[i * hw + px] : (cutout ? -1 : p_label[j * hw + px]);
                o_conf[i * hw + px] = keep ? conf[i * hw + px] : (cutout ? 0.0f : p_conf[j * hw + px]);
            }
    }
}
