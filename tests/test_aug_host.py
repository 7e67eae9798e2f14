"""CPU tests of the augmentation hand-off: the oracle restatement against the bundles recorded from the
live reference (tests/golden/aug_*.npz, cut_*.npz), Pillow's NEAREST table against Pillow itself, and the host-side
random draws of css_b200.aug (same Python / NumPy / torch RNG consumption as dataset_helpers/VOC.py:126-196, 511-535)."""
import random

import numpy as np
import pytest
import torch

from oracle import css_oracle as O
from tests.helpers import load_golden

AUG = ["aug_t2_scale", "aug_t2_flip", "aug_t3_city", "aug_t1_ori"]
CUT = ["cut_cutmix_2", "cut_cutmix_3", "cut_cutout_1", "cut_classmix_2"]
CUT_WORLD2 = "cut_cutmix_2_world2"          # generate_cut_gather_2 of the reference on a two-process group
CLASSMIX_WORLD2 = "cut_classmix_2_world2"   # the same with mode='classmix' and a different number of label values per image


def seed_all(seed):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def maps_of(g, prefix=""):
    labels = [g[f"{prefix}label{i}"] for i in range(2) if f"{prefix}label{i}" in g]
    confs = [g[f"{prefix}conf{i}"] for i in range(2) if f"{prefix}conf{i}" in g]
    return labels, confs


@pytest.mark.parametrize("name", AUG)
def test_oracle_aug_maps_vs_reference(name):
    g = load_golden(name)
    labels, confs = maps_of(g)
    ol, oc = O.aug_maps(labels, confs, g["geometry"], tuple(g["crop"]))
    ref_l, ref_c = maps_of(g, "out_")
    for a, b in zip(ol, ref_l):
        np.testing.assert_array_equal(a, b.astype(np.int64))
    for a, b in zip(oc, ref_c):
        np.testing.assert_array_equal(a, b)
    assert any((b == -1).any() for b in ref_l)                 # ignore pixels (incl. padding) are exercised


@pytest.mark.parametrize("name", CUT)
def test_oracle_cut_mix_vs_reference(name):
    g = load_golden(name)
    labels, confs = maps_of(g)
    labels = [l.astype(np.int64) for l in labels]
    sets = [[v for v in row if v != -100] for row in g["class_sets"]]
    o_img, o_lab, o_conf = O.cut_mix(g["image"], labels, confs, str(g["mode"]), boxes=g["boxes"], class_sets=sets)
    np.testing.assert_array_equal(o_img, g["out_image"])
    ref_l, ref_c = maps_of(g, "out_")
    for a, b in zip(o_lab, ref_l):
        np.testing.assert_array_equal(a, b.astype(np.int64))
    for a, b in zip(o_conf, ref_c):
        np.testing.assert_array_equal(a, b)


def test_oracle_cut_mix_two_ranks_vs_reference():
    """On more than one rank the reference's partner (i + 1) % batch_size is an image of RANK 0 and every rank draws a box for
    every gathered image (VOC.py:393-434 run on a two-process gloo group when the bundle was recorded)."""
    from css_b200 import aug
    g = load_golden(CUT_WORLD2)
    B, _, H, W = g["r0_image"].shape
    part = (g["r0_image"], [g["r0_label0"].astype(np.int64)], [g["r0_conf0"], g["r0_conf1"]])
    for rank in (0, 1):
        seed_all(int(g["seed"]))
        boxes = np.asarray([aug.draw_cut_box(H, W, 2) for _ in range(2 * B)])[rank * B:(rank + 1) * B]
        np.testing.assert_array_equal(boxes, g[f"r{rank}_boxes"])
        o = O.cut_mix(g[f"r{rank}_image"], [g[f"r{rank}_label0"].astype(np.int64)], [g[f"r{rank}_conf0"], g[f"r{rank}_conf1"]],
                      "cutmix", boxes=boxes, partner=part)
        np.testing.assert_array_equal(o[0], g[f"r{rank}_out_image"])
        np.testing.assert_array_equal(o[1][0], g[f"r{rank}_out_label0"].astype(np.int64))
        np.testing.assert_array_equal(o[2][0], g[f"r{rank}_out_conf0"])
        np.testing.assert_array_equal(o[2][1], g[f"r{rank}_out_conf1"])


def test_oracle_class_mix_two_ranks_vs_reference():
    """ClassMix on two ranks (VOC.py:423 -> :511-516): every rank draws one torch.randperm per GATHERED image, in gathered order.
    A rank needs only the NUMBER of label values present in the other ranks' images to stay in step with the generator -- the
    rule css_b200.aug follows (an all_gather of counts instead of label maps)."""
    import torch
    from css_b200 import aug
    g = load_golden(CLASSMIX_WORLD2)
    B = g["r0_image"].shape[0]
    part = (g["r0_image"], [g["r0_label0"].astype(np.int64)], [g["r0_conf0"], g["r0_conf1"]])
    counts = [len(np.unique(g[f"r{i // B}_label0"][i % B])) for i in range(2 * B)]
    assert len(set(counts)) > 2, "the bundle should exercise permutations of different lengths"
    for rank in (0, 1):
        seed_all(int(g["seed"]))
        sets = []
        for i in range(2 * B):
            if i // B == rank:
                sets.append(aug.draw_class_set(torch.from_numpy(g[f"r{rank}_label0"][i % B].astype(np.int64))))
            else:
                torch.randperm(counts[i])
        rec = [[v for v in row if v != -100] for row in g[f"r{rank}_class_sets"]]
        assert sets == rec
        o = O.cut_mix(g[f"r{rank}_image"], [g[f"r{rank}_label0"].astype(np.int64)], [g[f"r{rank}_conf0"], g[f"r{rank}_conf1"]],
                      "classmix", class_sets=sets, partner=part)
        np.testing.assert_array_equal(o[0], g[f"r{rank}_out_image"])
        np.testing.assert_array_equal(o[1][0], g[f"r{rank}_out_label0"].astype(np.int64))
        np.testing.assert_array_equal(o[2][0], g[f"r{rank}_out_conf0"])
        np.testing.assert_array_equal(o[2][1], g[f"r{rank}_out_conf1"])


def test_pil_nearest_table_matches_pillow():
    from PIL import Image
    rng = np.random.default_rng(0)
    sizes = [(321, 257), (321, 400), (500, 333), (769, 1538), (769, 385), (512, 768), (41, 26)]
    sizes += [(int(a), int(b)) for a, b in zip(rng.integers(8, 900, 25), rng.integers(8, 1600, 25))]
    for n_in, n_out in sizes:
        idx = np.arange(n_in)
        lo = Image.fromarray((idx & 255).astype(np.uint8)[None, :].repeat(2, 0))
        hi = Image.fromarray((idx >> 8).astype(np.uint8)[None, :].repeat(2, 0))
        got = np.asarray(lo.resize((n_out, 2), Image.NEAREST))[0].astype(np.int64) + \
            256 * np.asarray(hi.resize((n_out, 2), Image.NEAREST))[0].astype(np.int64)
        np.testing.assert_array_equal(O.pil_nearest_table(n_in, n_out), got, err_msg=f"{n_in}->{n_out}")


def test_byte_round_trips_are_exhaustive():
    k = np.arange(-1, 256)
    np.testing.assert_array_equal(O.label_to_byte(k), (k & 255).astype(np.uint8))
    back = O.byte_to_label(np.arange(256).astype(np.uint8))
    np.testing.assert_array_equal(back[:255], np.arange(255))
    assert back[255] == -1
    q = O.byte_to_conf(O.conf_to_byte(np.float32([0.0, 0.5, 0.999, 1.0])))
    np.testing.assert_array_equal(q, np.float32([0, 127, 254, 255]) / np.float32(255))


@pytest.mark.parametrize("name", AUG)
def test_host_geometry_and_image_path_replay_the_reference(name):
    """css_b200.aug's image path consumes the RNG streams like transform_* does: same seeds -> the reference's augmented
    image bit for bit, and the geometry the maps are replayed with."""
    import torchvision.transforms.functional as TF
    from css_b200 import aug
    g = load_golden(name)
    seed_all(int(g["seed"]))
    host = aug._unnormalise(torch.from_numpy(g["image"]))
    crop, scale = tuple(int(v) for v in g["crop"]), tuple(float(v) for v in g["scale"])
    out = [aug._augment_image(TF.to_pil_image(host[k]), crop, scale, bool(g["augmentation"])) for k in range(host.shape[0])]
    np.testing.assert_array_equal(np.asarray([o[1] for o in out]), g["geometry"])
    np.testing.assert_array_equal(torch.stack([o[0] for o in out]).numpy(), g["out_image"])


def test_host_cut_box_draws_replay_the_reference():
    from css_b200 import aug
    for name in ("cut_cutmix_2", "cut_cutout_1"):
        g = load_golden(name)
        seed_all(int(g["seed"]))
        B, _, H, W = g["image"].shape
        np.testing.assert_array_equal(np.asarray([aug.draw_cut_box(H, W, 2) for _ in range(B)]), g["boxes"])
    g = load_golden("cut_classmix_2")
    seed_all(int(g["seed"]))
    for i in range(g["image"].shape[0]):
        chosen = aug.draw_class_set(torch.from_numpy(g["label0"][i].astype(np.int64)))
        assert chosen == [v for v in g["class_sets"][i] if v != -100]


def test_aug_refuses_cpu_tensors():
    from css_b200 import aug
    with pytest.raises(RuntimeError, match="CUDA"):
        aug.transform_maps([torch.zeros(1, 4, 4)], [], np.zeros((1, 5), np.int32), (4, 4))
    with pytest.raises(RuntimeError, match="CUDA"):
        aug.batch_transform_2(torch.zeros(1, 3, 4, 4), torch.zeros(1, 4, 4), torch.zeros(1, 4, 4), torch.zeros(1, 4, 4))
    with pytest.raises(RuntimeError, match="CUDA"):
        aug.cut_mix(torch.zeros(1, 3, 4, 4), [torch.zeros(1, 4, 4, dtype=torch.int64)], [torch.zeros(1, 4, 4)], "cutmix",
                    boxes=[[0, 1, 0, 1]])
