"""css_grad_scatter through the C ABI against a NumPy scatter-add: the one-pass slab writer (per-image anchor list in shared
memory, list overflow, images smaller than a slab, duplicate and absent anchors) and the memset + atomics fallback (more than
32 k anchors, or a gradient buffer that is not 16-byte aligned).  Backward of loss.py:141-149 w.r.t. `rep`: dense [B2,256,h,w]."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
D = 256


def reference(anchor_px, grad_anchor, go, B2, h, w):
    out = np.zeros((B2, D, h * w), np.float64)
    for a, px in enumerate(anchor_px):
        if px >= 0:
            out[px // (h * w), :, px % (h * w)] += go * grad_anchor[a].astype(np.float64)
    return out.reshape(B2, D, h, w)


def run(anchor_px, grad_anchor, go, B2, h, w, misalign=False):
    from css_b200 import _lib
    from css_b200._lib import check, ptr, stream_ptr
    lib = _lib.load()
    dev = torch.device("cuda")
    px = torch.from_numpy(anchor_px.astype(np.int32)).to(dev)
    ga = torch.from_numpy(grad_anchor).to(dev)
    g0 = torch.full((), go, device=dev)
    n = B2 * D * h * w
    buf = torch.full((n + 4,), float("nan"), device=dev)           # poisoned: every element must be written
    out = buf[1:n + 1] if misalign else buf[:n]
    check(lib.css_grad_scatter(ptr(g0), ptr(px), ptr(ga), len(anchor_px), B2, D, h, w, out.data_ptr(), stream_ptr()), "css_grad_scatter")
    torch.cuda.synchronize()
    assert torch.isnan(buf[n + 1:]).all() and (misalign is False or torch.isnan(buf[0]))     # nothing written outside
    return out.cpu().numpy().reshape(B2, D, h, w)


CASES = {
    "voc_like": dict(B2=4, h=81, w=81, n=2000, spread="uniform"),
    "all_in_one_image_list_overflow": dict(B2=8, h=20, w=20, n=2048, spread="image3"),
    "image_smaller_than_a_slab": dict(B2=5, h=3, w=3, n=300, spread="uniform"),
    "single_image": dict(B2=1, h=33, w=29, n=700, spread="uniform"),
    "heavy_duplicates_and_absent": dict(B2=3, h=10, w=12, n=1500, spread="dups"),
    "more_than_32k_anchors_fallback": dict(B2=2, h=16, w=16, n=33000, spread="uniform"),
    "misaligned_buffer_fallback": dict(B2=2, h=9, w=7, n=400, spread="uniform", misalign=True),
}


@pytest.mark.parametrize("name", list(CASES))
def test_grad_scatter_matches_numpy(name):
    c = CASES[name]
    B2, h, w, n = c["B2"], c["h"], c["w"], c["n"]
    rng = np.random.default_rng(len(name) + n)
    N = B2 * h * w
    if c["spread"] == "uniform":
        px = rng.integers(0, N, n)
    elif c["spread"] == "image3":
        px = 3 * h * w + rng.integers(0, h * w, n)
    else:
        px = rng.choice(rng.integers(0, N, 7), n)                  # seven distinct pixels, hit hundreds of times each
        px[rng.random(n) < 0.3] = -1                               # slots of classes without a hard pixel
    ga = rng.standard_normal((n, D)).astype(np.float32)
    got = run(px, ga, 0.37, B2, h, w, c.get("misalign", False))
    ref = reference(px, ga, np.float32(0.37), B2, h, w)
    assert np.isfinite(got).all()
    assert np.array_equal(got != 0, ref != 0)                      # exactly the anchor pixels are non-zero
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=5e-5)     # fp32 atomics in arbitrary order vs float64
