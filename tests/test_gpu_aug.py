"""GPU parity of the augmentation hand-off: css_aug_index / css_aug_maps / css_cut_mix through the C ABI
against the bundles recorded from the live reference (batch_transform_*, generate_cut_gather_*), and against the oracle at
the VOC / CityScapes crop sizes.  Everything is byte / index work: the bar is exact equality."""
import numpy as np
import pytest
import torch

from oracle import css_oracle as O
from tests.helpers import load_golden
from tests.test_aug_host import AUG, CUT, maps_of, seed_all

pytestmark = pytest.mark.gpu


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t.to(dtype) if dtype is not None else t


@pytest.mark.parametrize("name", AUG)
def test_transform_maps_vs_reference(name):
    from css_b200 import aug
    g = load_golden(name)
    labels, confs = maps_of(g)
    ref_l, ref_c = maps_of(g, "out_")
    ol, oc = aug.transform_maps([dev(l) for l in labels], [dev(c) for c in confs], g["geometry"], tuple(g["crop"]))
    for a, b in zip(ol, ref_l):
        assert a.dtype == torch.int64
        np.testing.assert_array_equal(a.cpu().numpy(), b.astype(np.int64))
    for a, b in zip(oc, ref_c):
        np.testing.assert_array_equal(a.cpu().numpy(), b)
    # the second trip of a step feeds int64 labels with -1 (and already quantised confidences): same result from both dtypes
    as_int = [np.where(l == 255, -1, l).astype(np.int64) for l in labels]
    ol2, _ = aug.transform_maps([dev(l) for l in as_int], [], g["geometry"], tuple(g["crop"]))
    for a, b in zip(ol2, ref_l):
        np.testing.assert_array_equal(a.cpu().numpy(), b.astype(np.int64))


@pytest.mark.parametrize("name", AUG)
def test_batch_transform_drop_in_vs_reference(name):
    """Same seeds, same call as the reference's batch_transform_*: image, labels and confidences all identical."""
    from css_b200 import aug
    g = load_golden(name)
    labels, confs = maps_of(g)
    ref_l, ref_c = maps_of(g, "out_")
    kw = dict(crop_size=tuple(int(v) for v in g["crop"]), scale_size=tuple(float(v) for v in g["scale"]),
              augmentation=bool(g["augmentation"]))
    seed_all(int(g["seed"]))
    image = dev(g["image"])
    if int(g["maps"]) == 1:
        r = aug.batch_transform(image, dev(labels[0]), dev(confs[0]), **kw)
        got_l, got_c = [r[1]], [r[2]]
    elif int(g["maps"]) == 2:
        r = aug.batch_transform_2(image, dev(labels[0]), dev(confs[0]), dev(confs[1]), **kw)
        got_l, got_c = [r[1]], [r[2], r[3]]
    else:
        r = aug.batch_transform_3(image, dev(labels[0]), dev(labels[1]), dev(confs[0]), dev(confs[1]), **kw)
        got_l, got_c = [r[1], r[2]], [r[3], r[4]]
    assert r[0].is_cuda
    np.testing.assert_array_equal(r[0].cpu().numpy(), g["out_image"])
    for a, b in zip(got_l, ref_l):
        np.testing.assert_array_equal(a.cpu().numpy(), b.astype(np.int64))
    for a, b in zip(got_c, ref_c):
        np.testing.assert_array_equal(a.cpu().numpy(), b)


@pytest.mark.parametrize("name", CUT)
def test_generate_cut_gather_drop_in_vs_reference(name):
    from css_b200 import aug
    g = load_golden(name)
    labels, confs = maps_of(g)
    ref_l, ref_c = maps_of(g, "out_")
    seed_all(int(g["seed"]))
    image, mode = dev(g["image"]), str(g["mode"])
    L = [dev(l, torch.int64) for l in labels]
    Cf = [dev(c) for c in confs]
    if int(g["maps"]) == 1:
        r = aug.generate_cut_gather(image, L[0], Cf[0], mode=mode)
        got_l, got_c = [r[1]], [r[2]]
    elif int(g["maps"]) == 2:
        r = aug.generate_cut_gather_2(image, L[0], Cf[0], Cf[1], mode=mode)
        got_l, got_c = [r[1]], [r[2], r[3]]
    else:
        r = aug.generate_cut_gather_3(image, L[0], L[1], Cf[0], Cf[1], mode=mode)
        got_l, got_c = [r[1], r[2]], [r[3], r[4]]
    np.testing.assert_array_equal(r[0].cpu().numpy(), g["out_image"])
    for a, b in zip(got_l, ref_l):
        assert a.dtype == torch.int64
        np.testing.assert_array_equal(a.cpu().numpy(), b.astype(np.int64))
    for a, b in zip(got_c, ref_c):
        np.testing.assert_array_equal(a.cpu().numpy(), b)


def test_generate_cut_gather_none_and_bad_mode():
    from css_b200 import aug
    img, lab, c = torch.rand(2, 3, 8, 8).cuda(), torch.zeros(2, 8, 8).cuda(), torch.rand(2, 8, 8).cuda()
    r = aug.generate_cut_gather_2(img, lab, c, c, mode='none')
    assert r[1].dtype == torch.int64 and torch.equal(r[0], img)
    with pytest.raises(ValueError):
        aug.generate_cut_gather_2(img, lab, c, c, mode='mosaic')


@pytest.mark.parametrize("H,W,crop,B", [(321, 321, (321, 321), 8), (769, 769, (769, 769), 4), (200, 333, (160, 224), 3)])
def test_transform_maps_full_size_vs_oracle(H, W, crop, B):
    """Random geometry over the reference's scale range (0.5 .. 2.0: both padding and real crops), every map slot used."""
    from css_b200 import aug
    rng = np.random.default_rng(H + B)
    l0 = rng.integers(0, 21, (B, H, W)).astype(np.float32)
    l0[rng.random((B, H, W)) < 0.1] = 255
    l1 = rng.integers(-1, 19, (B, H, W)).astype(np.float32)
    l1[l1 < 0] = 255
    c0, c1 = rng.random((B, H, W), np.float32), rng.random((B, H, W), np.float32)
    c0[0, :3] = 1.0
    geo = np.zeros((B, 5), np.int32)
    for b in range(B):
        ratio = rng.uniform(0.5, 2.0)
        rh, rw = int(H * ratio), int(W * ratio)
        geo[b] = (rh, rw, rng.integers(0, max(rh, crop[0]) - crop[0] + 1), rng.integers(0, max(rw, crop[1]) - crop[1] + 1),
                  rng.integers(0, 2))
    ol, oc = aug.transform_maps([dev(l0), dev(l1)], [dev(c0), dev(c1)], geo, crop)
    rl, rc = O.aug_maps([l0, l1], [c0, c1], geo, crop)
    for a, b in zip(ol + oc, rl + rc):
        np.testing.assert_array_equal(a.cpu().numpy(), b)


def test_cut_mix_full_size_vs_oracle_with_partner_batch():
    """321 x 321, B = 8, partners taken from ANOTHER batch (what ranks > 0 do with rank 0's broadcast batch)."""
    from css_b200 import aug
    rng = np.random.default_rng(3)
    B, H, W = 8, 321, 321

    def batch():
        return (rng.standard_normal((B, 3, H, W), np.float32), [rng.integers(-1, 21, (B, H, W)), rng.integers(-1, 21, (B, H, W))],
                [rng.random((B, H, W), np.float32), rng.random((B, H, W), np.float32)])
    own, par = batch(), batch()
    np.random.seed(5)
    boxes = np.asarray([aug.draw_cut_box(H, W, 2) for _ in range(B)])
    to_dev = lambda t: (dev(t[0]), [dev(x) for x in t[1]], [dev(x) for x in t[2]])     # noqa: E731
    for mode, sets in (("cutmix", None), ("classmix", [list(rng.choice(np.arange(-1, 21), 10, replace=False)) for _ in range(B)])):
        d_own, d_par = to_dev(own), to_dev(par)
        o = aug.cut_mix(d_own[0], d_own[1], d_own[2], mode, boxes=boxes, class_sets=sets, partner=d_par)
        r = O.cut_mix(own[0], own[1], own[2], mode, boxes=boxes, class_sets=sets, partner=par)
        np.testing.assert_array_equal(o[0].cpu().numpy(), r[0])
        for a, b in zip(o[1] + o[2], r[1] + r[2]):
            np.testing.assert_array_equal(a.cpu().numpy(), b)


def test_model_mix_shell_with_gpu_aug_hooks():
    """Model_mix wired to css_b200.aug (install(gpu_aug=True) does this): the shell's output equals the three augmentation
    calls of ddp_model.py:121-135 made by hand on the stage-1/2 outputs under the same seeds, with the reference's dtypes."""
    import css_b200
    from css_b200 import aug, models

    class Stub(torch.nn.Module):
        def __init__(self, outs):
            super().__init__()
            self.outs, self.i = outs, 0

        def forward(self, x):
            o = self.outs[self.i % len(self.outs)]
            self.i += 1
            return o

    g = load_golden("stage12_mix_c21")
    B, C, H, W, temp = int(g["B"]), int(g["C"]), int(g["H"]), int(g["W"]), float(g["temp"])
    saved = dict(vars(models.hooks))
    try:
        models.hooks.network_factory = lambda enc, **kw: torch.nn.Conv2d(1, 1, 1)
        for n in ("batch_transform", "batch_transform_2", "batch_transform_3", "generate_cut_gather", "generate_cut_gather_2",
                  "generate_cut_gather_3"):
            setattr(models.hooks, n, getattr(aug, n))
        cfg = {"Dataset": {"crop_size": (H, W), "scale_size": (0.5, 1.5), "mix_mode": "cutmix"}}
        m = models.Model_mix(None, num_classes=C, output_dim=256, config=cfg, temp=temp).cuda()
        rep_all, pred_all = dev(g["rep_all"]), dev(g["pred_all"])
        m.ema_model = Stub([(dev(g["pred_u"]), dev(g["rep_u"]))])
        m.model = Stub([(pred_all[:B], rep_all[:B]), (pred_all[B:], rep_all[B:])])
        img = torch.rand(B, 3, H, W, generator=torch.Generator().manual_seed(1)).cuda()
        protos = dev(g["prototypes"])
        seed_all(77)
        r = m(img, img, protos)
        seed_all(77)
        o = css_b200.ops.pseudo_labels(dev(g["rep_u"]), dev(g["pred_u"]), protos, temp, (H, W), fuse="mix")
        a = aug.batch_transform_2(img, o["fused"], o["conf_cls"], o["conf_rep"], crop_size=(H, W), scale_size=(0.5, 1.5), augmentation=False)
        a = aug.generate_cut_gather_2(*a, mode="cutmix")
        a = aug.batch_transform_2(*a, crop_size=(H, W), scale_size=(1.0, 1.0), augmentation=True)
        assert r[2].dtype == torch.int64 and r[3].dtype == torch.float32 and r[2].shape == (B, H, W)
        assert torch.equal(r[2], a[1]) and torch.equal(r[3], a[2]) and torch.equal(r[4], a[3])
        lab = r[2].cpu().numpy()
        assert lab.min() >= -1 and lab.max() < C
        q = r[3].cpu().numpy() * np.float32(255)
        np.testing.assert_allclose(q, np.round(q), atol=1e-4)         # confidences come back as multiples of 1/255
    finally:
        for k, v in saved.items():
            setattr(models.hooks, k, v)
