"""The plain-C restatement of the path's byte / integer / index work (oracle/css_oracle_int.c, gcc) against the numpy oracle,
the golden bundles recorded from the live reference, and Pillow: two independent restatements have to agree exactly."""
import ctypes

import numpy as np
import pytest

from oracle import build_c
from oracle import css_oracle as O
from tests.helpers import load_golden
from tests.test_aug_host import AUG, maps_of

I64P = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
I32P = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
F32P = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
ci, cf = ctypes.c_int, ctypes.c_float


@pytest.fixture(scope="module")
def lib():
    L = ctypes.CDLL(build_c.build())
    L.orc_pil_nearest_table.argtypes = [ci, ci, I64P]
    L.orc_aug_label.argtypes = [I64P, ci, ci, ci, ci, ci, ci, ci, ci, ci, I64P]
    L.orc_aug_conf.argtypes = [F32P, ci, ci, ci, ci, ci, ci, ci, ci, ci, F32P]
    L.orc_mix_fuse.argtypes = [I64P, I64P, ctypes.c_int64, F32P]
    L.orc_select.argtypes = [F32P, F32P, F32P, cf, ci, ci, ci, ci, I32P, I32P, I32P, I32P]
    L.orc_threshold_glue.argtypes = [I64P, I64P, F32P, cf, ci, ci, ci, ci, ci, ci, ci, F32P, F32P]
    L.orc_cut_mix_boxes.argtypes = [F32P, I64P, F32P, F32P, I64P, F32P, I32P, ci, ci, ci, ci, ci, F32P, I64P, F32P]
    for f in ("orc_pil_nearest_table", "orc_aug_label", "orc_aug_conf", "orc_mix_fuse", "orc_select", "orc_threshold_glue",
              "orc_cut_mix_boxes"):
        getattr(L, f).restype = None
    return L


def c(a, dt):
    return np.ascontiguousarray(np.asarray(a), dtype=dt)


def test_c_nearest_table_vs_numpy_and_pillow(lib):
    from PIL import Image
    rng = np.random.default_rng(1)
    for n_in, n_out in [(321, 257), (500, 333), (769, 1538), (41, 26)] + [tuple(int(v) for v in rng.integers(8, 1500, 2)) for _ in range(30)]:
        out = np.empty(n_out, np.int64)
        lib.orc_pil_nearest_table(n_in, n_out, out)
        np.testing.assert_array_equal(out, O.pil_nearest_table(n_in, n_out))
        idx = np.arange(n_in)
        lo = Image.fromarray((idx & 255).astype(np.uint8)[None, :].repeat(2, 0))
        hi = Image.fromarray((idx >> 8).astype(np.uint8)[None, :].repeat(2, 0))
        pil = np.asarray(lo.resize((n_out, 2), Image.NEAREST))[0].astype(np.int64) + \
            256 * np.asarray(hi.resize((n_out, 2), Image.NEAREST))[0].astype(np.int64)
        np.testing.assert_array_equal(out, pil)


@pytest.mark.parametrize("name", AUG)
def test_c_aug_maps_vs_reference_bundles(lib, name):
    g = load_golden(name)
    labels, confs = maps_of(g)
    ref_l, ref_c = maps_of(g, "out_")
    ch, cw = (int(v) for v in g["crop"])
    for b, (rh, rw, top, left, flip) in enumerate(g["geometry"].tolist()):
        for m, r in zip(labels, ref_l):
            H, W = m.shape[1:]
            out = np.empty((ch, cw), np.int64)
            lib.orc_aug_label(c(m[b], np.int64), H, W, rh, rw, top, left, flip, ch, cw, out)
            np.testing.assert_array_equal(out, r[b].astype(np.int64))
        for m, r in zip(confs, ref_c):
            H, W = m.shape[1:]
            out = np.empty((ch, cw), np.float32)
            lib.orc_aug_conf(c(m[b], np.float32), H, W, rh, rw, top, left, flip, ch, cw, out)
            np.testing.assert_array_equal(out, r[b])


def test_c_mix_fuse_and_glue_vs_reference_bundles(lib):
    g = load_golden("stage12_cross_c19")                          # the reference's own two label maps
    lc, lr = c(g["label_cls"], np.int64), c(g["label_rep"], np.int64)
    out = np.empty(lc.shape, np.float32)
    lib.orc_mix_fuse(lc, lr, lc.size, out)
    np.testing.assert_array_equal(out, O.mix_fuse(lc, lr, int(g["C"])))
    g = load_golden("stage12_mix_c21")                            # the reference's fused map from the oracle's label maps
    H, W = int(g["H"]), int(g["W"])
    _, lc = O.cls_pseudo_label(g["pred_u"], (H, W))
    _, lr, _ = O.rep_pseudo_label(g["rep_u"], g["prototypes"], float(g["temp"]), (H, W))
    lc, lr = c(lc, np.int64), c(lr, np.int64)
    out = np.empty(lc.shape, np.float32)
    lib.orc_mix_fuse(lc, lr, lc.size, out)
    np.testing.assert_array_equal(out, O.mix_fuse(lc, lr, int(g["C"])))
    assert (out != g["fused"]).mean() < 1e-3                      # near-tie pixels only (see test_oracle_golden.py)
    for name, mode in (("glue_mix", 1), ("glue_cross", 0), ("glue_ori", 0)):
        g = load_golden(name)
        ll, lu, conf = c(g["train_l_label"], np.int64), c(g["u_label"], np.int64), c(g["conf"], np.float32)
        B, H, W = ll.shape
        C, (h, w) = int(g["C"]), g["label_all"].shape[2:]
        la, ma = np.empty((2 * B, C, h, w), np.float32), np.empty((2 * B, 1, h, w), np.float32)
        lib.orc_threshold_glue(ll, lu, conf, float(g["weak"]), mode, B, C, H, W, h, w, la, ma)
        np.testing.assert_array_equal(la, g["label_all"])
        np.testing.assert_array_equal(ma, g["mask_all"])


@pytest.mark.parametrize("B2,C,h,w,multi_hot", [(2, 21, 13, 11, False), (3, 5, 7, 9, True), (1, 32, 4, 4, False)])
def test_c_selection_vs_numpy_row_major_lists(lib, B2, C, h, w, multi_hot):
    rng = np.random.default_rng(B2 * 100 + C)
    cls = rng.integers(0, C, (B2, h, w))
    label = np.zeros((B2, C, h, w), np.float32)
    np.put_along_axis(label, cls[:, None], 1.0, axis=1)
    if multi_hot:
        label[:, 1] = np.maximum(label[:, 1], (rng.random((B2, h, w)) < 0.3).astype(np.float32))
    mask = (rng.random((B2, 1, h, w)) < 0.8).astype(np.float32)
    prob = rng.random((B2, C, h, w)).astype(np.float32)
    strong = 0.6
    N = B2 * h * w
    vl, hl = np.full(C * N, -1, np.int32), np.full(C * N, -1, np.int32)
    nv, nh = np.zeros(C, np.int32), np.zeros(C, np.int32)
    lib.orc_select(label, mask, prob, strong, B2, C, h, w, vl, hl, nv, nh)
    valid = (label * mask) != 0                                   # loss.py:80,99,111
    hard = valid & (prob < np.float32(strong))
    for k in range(C):
        ref_v = np.flatnonzero(valid[:, k].reshape(-1))
        ref_h = np.flatnonzero(hard[:, k].reshape(-1))
        assert nv[k] == len(ref_v) and nh[k] == len(ref_h)
        np.testing.assert_array_equal(vl[k * N:k * N + nv[k]], ref_v)
        np.testing.assert_array_equal(hl[k * N:k * N + nh[k]], ref_h)


@pytest.mark.parametrize("name", ["cut_cutmix_2", "cut_cutout_1"])
def test_c_cut_mix_vs_reference_bundles(lib, name):
    g = load_golden(name)
    image, label, conf = c(g["image"], np.float32), c(g["label0"], np.int64), c(g["conf0"], np.float32)
    B, CH, H, W = image.shape
    o_img, o_lab, o_conf = np.empty_like(image), np.empty_like(label), np.empty_like(conf)
    lib.orc_cut_mix_boxes(image, label, conf, image, label, conf, c(g["boxes"], np.int32), int(str(g["mode"]) == "cutout"), B, CH, H, W,
                          o_img, o_lab, o_conf)
    np.testing.assert_array_equal(o_img, g["out_image"])
    np.testing.assert_array_equal(o_lab, g["out_label0"].astype(np.int64))
    np.testing.assert_array_equal(o_conf, g["out_conf0"])
