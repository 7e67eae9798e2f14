"""Generate the golden bundles under tests/golden/ by running the LIVE, UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):   python -m oracle.gen_golden
The reference holds no golden vectors of its own, so these bundles -- inputs, recorded
RNG draws, outputs, autograd gradients of the reference itself -- are what pins oracle/css_oracle.py.
Each bundle is a small .npz; the GPU box only ever reads the committed files.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_harness import load_reference, DrawRecorder, StubNet  # noqa: E402
from css_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy().copy()      # copy: .numpy() aliases tensors that are updated in place later


def loss_case(name, *, B2, C, h, w, Q, Nn, temp, strong, alpha, seed, strategy, protos, steps=1, tweak=None,
              block=8):
    """Contrast_Loss.forward + backward of the reference (loss.py:66-149) on synthetic inputs, `steps` times with
    the prototypes carried over (first-touch branch, then EMA branch)."""
    L, M, U = load_reference()
    crit = L.Contrast_Loss(num_queries=Q, num_negatives=Nn, temp=temp, strong_threshold=strong, alpha=alpha)
    out = dict(B2=B2, C=C, h=h, w=w, Q=Q, Nn=Nn, temp=temp, strong=strong, alpha=alpha, steps=steps, block=block)
    prototypes = protos.clone()
    for s in range(steps):
        d = synth.student_batch(B2, C, h, w, seed=seed + s, strategy=strategy, block=block)
        if strategy in ("mix", "cross"):
            # prob_all as Model_mix/Model_cross compute it from the *incoming* prototypes (ddp_model.py:147-154)
            xn = F.normalize(d["rep"].permute(0, 2, 3, 1), dim=-1).reshape(-1, d["rep"].shape[1])
            pn = F.normalize(prototypes, dim=-1).permute(1, 0)
            d["prob"] = F.softmax(torch.mm(xn, pn).reshape(B2, h, w, C).permute(0, 3, 1, 2) / temp, dim=1).contiguous()
        if tweak is not None:
            tweak(d, s)
        rep = d["rep"].clone().requires_grad_(True)
        torch.manual_seed(seed + 1000 + s)
        np.random.seed(seed + 2000 + s)
        rec = DrawRecorder()
        proto_in = prototypes.clone()
        with rec.recording():
            loss = crit(rep, d["label"], d["mask"], d["prob"], prototypes)
        loss.backward()
        out.update({
            f"s{s}_rep": _np(d["rep"]), f"s{s}_label": _np(d["label"]).astype(np.uint8),
            f"s{s}_mask": _np(d["mask"]).astype(np.uint8), f"s{s}_prob": _np(d["prob"]),
            f"s{s}_proto_in": _np(proto_in), f"s{s}_proto_out": _np(prototypes),
            f"s{s}_loss": np.float32(loss.item()), f"s{s}_grad": _np(rep.grad),
            f"s{s}_torch_seed": seed + 1000 + s, f"s{s}_numpy_seed": seed + 2000 + s,
            f"s{s}_n_scored": len(rec.anchor_idx),
        })
        for k, (a, c, n) in enumerate(zip(rec.anchor_idx, rec.samp_class, rec.neg_idx)):
            out[f"s{s}_anchor_idx_{k}"] = a.astype(np.int32)
            out[f"s{s}_samp_class_{k}"] = c.astype(np.uint8)
            out[f"s{s}_neg_idx_{k}"] = n.astype(np.int32)
        print(f"  {name} step {s}: loss={loss.item():.6f} scored={len(rec.anchor_idx)} |grad|={rep.grad.abs().sum().item():.4f}")
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def _passthrough2(images, labels, logits_1=None, logits_2=None, **kw):
    return images, labels, logits_1, logits_2


def _passthrough3(images, l1, l2, logits_1=None, logits_2=None, **kw):
    return images, l1, l2, logits_1, logits_2


def _passthrough1(images, labels, logits=None, **kw):
    return images, labels, logits


def stage12_case(name, *, kind, B, C, h, w, H, W, temp, seed, zero_rows=()):
    """Model_mix / Model_cross / Model_ori_pseudo.forward of the reference with a stub network and pass-through
    augmentation, so exactly ddp_model.py:101-118,147-154 (mix), :186-199,230-237 (cross), :34-37 (ori) execute."""
    L, M, U = load_reference()
    import torchvision.models as models
    cfg = {"Dataset": {"crop_size": (H, W), "scale_size": (1.0, 1.0), "mix_mode": "none"}}
    t = synth.teacher_batch(B, C, h, w, seed=seed)
    s = synth.student_batch(2 * B, C, h, w, seed=seed + 1)
    protos = 0.5 * t["centers"] + 0.3 * synth.warm_prototypes(C, seed=seed)
    for r in zero_rows:
        protos[r] = 0
    pred_l_t = synth.logits_for(synth.class_map(B, C, h, w, synth._gen(seed + 5)), C, synth._gen(seed + 6))
    rep_l_t = torch.randn(B, 256, h, w, generator=synth._gen(seed + 8))
    img_l = torch.zeros(B, 3, H, W)
    img_u = torch.zeros(B, 3, H, W)
    ema = StubNet([(pred_l_t, rep_l_t), (t["pred_u"], t["rep_u"])])
    stu = StubNet([(s["logits"][:B], s["rep"][:B]), (s["logits"][B:], s["rep"][B:])])
    saved = (M.batch_transform, M.batch_transform_2, M.batch_transform_3,
             M.generate_cut_gather, M.generate_cut_gather_2, M.generate_cut_gather_3)
    M.batch_transform, M.batch_transform_2, M.batch_transform_3 = _passthrough1, _passthrough2, _passthrough3
    M.generate_cut_gather = lambda a, b, c, mode=None: (a, b, c)
    M.generate_cut_gather_2 = lambda a, b, c, d, mode=None: (a, b, c, d)
    M.generate_cut_gather_3 = lambda a, b, c, d, e, mode=None: (a, b, c, d, e)
    try:
        out = dict(kind=kind, B=B, C=C, h=h, w=w, H=H, W=W, temp=temp,
                   rep_u=_np(t["rep_u"]), pred_u=_np(t["pred_u"]), prototypes=_np(protos),
                   rep_all=_np(s["rep"]), pred_all=_np(s["logits"]))
        if kind == "mix":
            m = M.Model_mix(models.resnet18(), num_classes=C, output_dim=256, config=cfg, temp=temp)
            m.model, m.ema_model = stu, ema
            r = m(img_l, img_u, protos)
            out.update(fused=_np(r[2]), conf_cls=_np(r[3]), conf_rep=_np(r[4]), prob_all=_np(r[6]))
        elif kind == "cross":
            m = M.Model_cross(models.resnet18(), num_classes=C, output_dim=256, config=cfg, temp=temp)
            m.model, m.ema_model = stu, ema
            r = m(img_l, img_u, protos)
            out.update(label_cls=_np(r[2]), label_rep=_np(r[3]), conf_cls=_np(r[4]), conf_rep=_np(r[5]),
                       prob_all=_np(r[7]))
        else:
            m = M.Model_ori_pseudo(models.resnet18(), num_classes=C, output_dim=256, config=cfg)
            m.model, m.ema_model = stu, StubNet([(t["pred_u"], t["rep_u"])])
            r = m(img_l, img_u)
            out.update(label_cls=_np(r[2]), conf_cls=_np(r[3]), pred_u_large_raw=_np(r[6]))
        # the up-sampled similarity map itself (ddp_model.py:111), recomputed with the same torch calls, so the
        # oracle's bilinear restatement and the near-tie margins can be checked directly
        xn = F.normalize(t["rep_u"].permute(0, 2, 3, 1), dim=-1).reshape(-1, 256)
        sim = torch.mm(xn, F.normalize(protos, dim=-1).permute(1, 0)).reshape(B, h, w, C).permute(0, 3, 1, 2)
        out.update(sim=_np(sim), sim_large=_np(F.interpolate(sim, size=(H, W), mode="bilinear", align_corners=True)))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print(f"  {name}: ok ({kind})")
    finally:
        (M.batch_transform, M.batch_transform_2, M.batch_transform_3,
         M.generate_cut_gather, M.generate_cut_gather_2, M.generate_cut_gather_3) = saved


def glue_case(name, *, strategy, B, C, H, W, h, w, weak, seed):
    """The no_grad glue block of train(): mix_label.py:175-183 / cross_label.py:178-185 / ori_pseudo.py:171-178.
    The scripts cannot be imported (`shutup` missing, and the block is inline in train()), so the same statements
    are executed here with the reference's own label_onehot / label_onehot_2 (utils.py:116-136)."""
    L, M, U = load_reference()
    g = synth._gen(seed)
    train_l_label = synth.class_map(B, C, H, W, g, ignore_frac=0.1, block=16)
    u_label = synth.class_map(B, C, H, W, g, ignore_frac=0.2, block=16)
    conf = torch.rand(B, H, W, generator=g)
    conf = torch.floor(conf * 255) / 255           # 8-bit quantised as after the PIL round trip (VOC.py:289-290)
    with torch.no_grad():
        u_mask = conf.ge(weak).float()
        mask_all = torch.cat(((train_l_label.unsqueeze(1) >= 0).float(), u_mask.unsqueeze(1)))
        mask_all = F.interpolate(mask_all, size=(h, w), mode="nearest")
        label_l = F.interpolate(U.label_onehot(train_l_label, C), size=(h, w), mode="nearest")
        if strategy == "mix":
            label_u = F.interpolate(U.label_onehot_2(u_label, C), size=(h, w), mode="nearest")[:, 1:]
        else:
            label_u = F.interpolate(U.label_onehot(u_label, C), size=(h, w), mode="nearest")
        label_all = torch.cat((label_l, label_u))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), strategy=strategy, C=C, weak=weak, h=h, w=w,
                        train_l_label=_np(train_l_label).astype(np.int16), u_label=_np(u_label).astype(np.int16),
                        conf=_np(conf), mask_all=_np(mask_all).astype(np.uint8),
                        label_all=_np(label_all).astype(np.uint8))
    print(f"  {name}: ok")


def atl_case(name, *, B, C, H, W, thr, seed):
    """Attention_Threshold_Loss forward + backward of the reference (loss.py:48-64)."""
    L, M, U = load_reference()
    g = synth._gen(seed)
    cls = synth.class_map(B, C, H, W, g, ignore_frac=0.15, block=6)
    pred = synth.logits_for(cls, C, g).requires_grad_(True)
    conf = torch.floor(torch.rand(B, H, W, generator=g) * 255) / 255
    cls[0, :2] = -1
    crit = L.Attention_Threshold_Loss(strong_threshold=thr)
    loss = crit(pred, cls, conf)
    (loss * 1.7).backward()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), C=C, thr=thr, pred=_np(pred), label=_np(cls).astype(np.int16), conf=_np(conf),
                        loss=np.float32(loss.item()), grad=_np(pred.grad), grad_scale=np.float32(1.7))
    print(f"  {name}: loss={loss.item():.6f}")


def _seed_all(seed):
    import random
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def _aug_inputs(B, C, H, W, seed, int_labels):
    import torchvision.transforms.functional as TF
    g = synth._gen(seed)
    image = TF.normalize(torch.rand(B, 3, H, W, generator=g), mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])
    cls = synth.class_map(B, C, H, W, g, ignore_frac=0.1, block=5)          # -1 = ignore
    cls2 = synth.class_map(B, C, H, W, g, ignore_frac=0.1, block=7)
    if not int_labels:                                                        # the fused float map of Model_mix: 255 = ignore
        cls = torch.where(cls < 0, torch.full_like(cls, 255), cls).float()
        cls2 = torch.where(cls2 < 0, torch.full_like(cls2, 255), cls2).float()
    return image, cls, cls2, torch.rand(B, H, W, generator=g), torch.rand(B, H, W, generator=g)


def aug_case(name, *, maps, B, C, H, W, crop, scale, augmentation, seed, int_labels=False):
    """batch_transform / _2 / _3 of the reference (dataset_helpers/VOC.py:312-352) under fixed Python / NumPy / torch seeds.
    The bundle also stores the geometry css_b200.aug draws from the same seeds (checked here against the reference's
    outputs through the oracle), so the GPU-side replay can be tested without PIL."""
    load_reference()
    import generalframeworks.dataset_helpers.VOC as V
    import torchvision.transforms.functional as TF
    from css_b200 import aug
    from oracle import css_oracle as O
    image, l1, l2, c1, c2 = _aug_inputs(B, C, H, W, seed, int_labels)
    _seed_all(seed)
    if maps == 1:
        r = V.batch_transform(image, l1, c1, crop_size=crop, scale_size=scale, augmentation=augmentation)
        labels, confs, ref_l, ref_c = [l1], [c1], [r[1]], [r[2]]
    elif maps == 2:
        r = V.batch_transform_2(image, l1, c1, c2, crop_size=crop, scale_size=scale, augmentation=augmentation)
        labels, confs, ref_l, ref_c = [l1], [c1, c2], [r[1]], [r[2], r[3]]
    else:
        r = V.batch_transform_3(image, l1, l2, c1, c2, crop_size=crop, scale_size=scale, augmentation=augmentation)
        labels, confs, ref_l, ref_c = [l1, l2], [c1, c2], [r[1], r[2]], [r[3], r[4]]
    _seed_all(seed)
    host = aug._unnormalise(image)
    mine = [aug._augment_image(TF.to_pil_image(host[k]), crop, scale, augmentation) for k in range(B)]
    geometry = np.asarray([m[1] for m in mine], np.int32)
    assert torch.equal(torch.stack([m[0] for m in mine]), r[0]), "image path / RNG order drifted from the reference"
    ol, oc = O.aug_maps([_np(t) for t in labels], [_np(t) for t in confs], geometry, crop)
    assert all(np.array_equal(a, _np(b)) for a, b in zip(ol, ref_l)) and all(np.array_equal(a, _np(b)) for a, b in zip(oc, ref_c))
    d = dict(maps=maps, seed=seed, crop=np.asarray(crop), scale=np.asarray(scale, np.float64), augmentation=augmentation,
             image=_np(image), geometry=geometry, out_image=_np(r[0]))
    for i, t in enumerate(labels):
        d[f"label{i}"] = _np(t)
        d[f"out_label{i}"] = _np(ref_l[i]).astype(np.int16)
    for i, t in enumerate(confs):
        d[f"conf{i}"] = _np(t)
        d[f"out_conf{i}"] = _np(ref_c[i])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(f"  {name}: geometry {geometry.tolist()}")


def cut_case(name, *, maps, mode, B, C, H, W, seed):
    """generate_cut_gather / _2 / _3 of the reference (VOC.py:354-477) on a single-process group."""
    load_reference()
    import generalframeworks.dataset_helpers.VOC as V
    from css_b200 import aug
    from oracle import css_oracle as O
    image, l1, l2, c1, c2 = _aug_inputs(B, C, H, W, seed, True)
    _seed_all(seed)
    if maps == 1:
        r = V.generate_cut_gather(image.clone(), l1.clone(), c1.clone(), mode=mode)
        labels, confs, ref_l, ref_c = [l1], [c1], [r[1]], [r[2]]
    elif maps == 2:
        r = V.generate_cut_gather_2(image.clone(), l1.clone(), c1.clone(), c2.clone(), mode=mode)
        labels, confs, ref_l, ref_c = [l1], [c1, c2], [r[1]], [r[2], r[3]]
    else:
        r = V.generate_cut_gather_3(image.clone(), l1.clone(), l2.clone(), c1.clone(), c2.clone(), mode=mode)
        labels, confs, ref_l, ref_c = [l1, l2], [c1, c2], [r[1], r[2]], [r[3], r[4]]
    _seed_all(seed)
    boxes = np.zeros((B, 4), np.int32)
    sets = np.full((B, 64), -100, np.int32)                                  # ragged class sets, padded with -100
    if mode == "classmix":
        for i in range(B):
            chosen = aug.draw_class_set(labels[0][i])
            sets[i, :len(chosen)] = chosen
    else:
        boxes = np.asarray([aug.draw_cut_box(H, W, 2) for _ in range(B)], np.int32)
    o_img, o_lab, o_conf = O.cut_mix(_np(image), [_np(t) for t in labels], [_np(t) for t in confs], mode, boxes=boxes,
                                     class_sets=[[v for v in row if v != -100] for row in sets])
    assert np.array_equal(o_img, _np(r[0])), "cut_mix restatement drifted from the reference"
    assert all(np.array_equal(a, _np(b)) for a, b in zip(o_lab, ref_l)) and all(np.array_equal(a, _np(b)) for a, b in zip(o_conf, ref_c))
    d = dict(maps=maps, mode=mode, seed=seed, image=_np(image), boxes=boxes, class_sets=sets, out_image=_np(r[0]))
    for i, t in enumerate(labels):
        d[f"label{i}"] = _np(t).astype(np.int16)
        d[f"out_label{i}"] = _np(ref_l[i]).astype(np.int16)
    for i, t in enumerate(confs):
        d[f"conf{i}"] = _np(t)
        d[f"out_conf{i}"] = _np(ref_c[i])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(f"  {name}: boxes {boxes.tolist()}")


def _cut_world2_worker(rank, port, out, mode):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=2)
    load_reference()
    import generalframeworks.dataset_helpers.VOC as V
    image, l1, _, c1, c2 = _aug_inputs(3, 21, 33, 37, 300 + rank, True)
    if mode == "classmix":      # a different number of label values per image and rank: the permutation draws differ in length
        for i in range(l1.shape[0]):
            l1[i] = torch.where(l1[i] >= 0, l1[i] % (4 + 3 * i + 7 * rank), l1[i])
    _seed_all(31)                                     # the scripts seed every rank alike (args.seed)
    r = V.generate_cut_gather_2(image.clone(), l1.clone(), c1.clone(), c2.clone(), mode=mode)
    out[rank] = dict(image=_np(image), label0=_np(l1).astype(np.int16), conf0=_np(c1), conf1=_np(c2), out_image=_np(r[0]),
                     out_label0=_np(r[1]).astype(np.int16), out_conf0=_np(r[2]), out_conf1=_np(r[3]))
    dist.destroy_process_group()


def cut_world2_case(name, mode="cutmix", port=29641):
    """generate_cut_gather_2 of the reference on a TWO-process gloo group (VOC.py:393-434): pins that the partner index
    (i + 1) % batch_size lands in rank 0's slice of the gathered batch and that every rank draws boxes (cutmix) / class
    permutations (classmix) for all gathered images, in gathered order."""
    import torch.multiprocessing as mp
    from css_b200 import aug
    from oracle import css_oracle as O
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_cut_world2_worker, args=(port, out, mode), nprocs=2, join=True)
    d = dict(seed=31, mode=mode)
    B, _, H, W = out[0]["image"].shape
    for rank in (0, 1):
        _seed_all(31)
        own, r0 = out[rank], out[0]
        boxes = np.zeros((B, 4), np.int32)
        sets = np.full((B, 64), -100, np.int32)
        if mode == "classmix":
            for i in range(2 * B):
                src = out[i // B]["label0"][i % B].astype(np.int64)
                chosen = aug.draw_class_set(torch.from_numpy(src))          # the reference's draw for gathered image i
                if i // B == rank:
                    sets[i % B, :len(chosen)] = chosen
        else:
            boxes = np.asarray([aug.draw_cut_box(H, W, 2) for _ in range(2 * B)], np.int32)[rank * B:(rank + 1) * B]
        o = O.cut_mix(own["image"], [own["label0"].astype(np.int64)], [own["conf0"], own["conf1"]], mode, boxes=boxes,
                      class_sets=[[v for v in row if v != -100] for row in sets],
                      partner=(r0["image"], [r0["label0"].astype(np.int64)], [r0["conf0"], r0["conf1"]]))
        assert np.array_equal(o[0], own["out_image"]) and np.array_equal(o[1][0], own["out_label0"].astype(np.int64))
        assert np.array_equal(o[2][0], own["out_conf0"]) and np.array_equal(o[2][1], own["out_conf1"])
        d[f"r{rank}_boxes"] = boxes
        d[f"r{rank}_class_sets"] = sets
        for k, v in own.items():
            d[f"r{rank}_{k}"] = v
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(f"  {name}: ok")


def main():
    os.makedirs(OUT, exist_ok=True)
    D = 256
    zeros = lambda C: torch.zeros(C, D)  # noqa: E731

    print("loss cases")
    # first-touch then EMA (two steps, prototypes carried over), ori_pseudo-style prob
    loss_case("loss_ori_c5", B2=2, C=5, h=9, w=9, Q=8, Nn=16, temp=0.5, strong=0.97, alpha=0.99, seed=11,
              strategy="ori", protos=zeros(5), steps=2)
    # VOC class count, ragged map, mix-style labels (all-zero rows on ignored pixels), prob from prototypes
    loss_case("loss_mix_c21", B2=2, C=21, h=13, w=11, Q=16, Nn=32, temp=0.5, strong=0.8, alpha=0.99, seed=23,
              strategy="mix", protos=synth.warm_prototypes(21, seed=23, zero_rows=(3,)), steps=2, block=2)

    # a present class without any hard pixel (skipped at loss.py:125-130 but counted in V), an absent class,
    # and a lower strong threshold so that hard != valid
    def tweak_nohard(d, s):
        d["prob"][:, 2] = 0.99                       # class 2: valid but never hard
        d["mask"][d["cls"].unsqueeze(1) == 4] = 0    # class 4: absent
        d["prob"][:, 0] = torch.where(d["prob"][:, 0] > 0.5, torch.full_like(d["prob"][:, 0], 0.98), d["prob"][:, 0])
    loss_case("loss_nohard_c7", B2=2, C=7, h=10, w=12, Q=12, Nn=24, temp=0.3, strong=0.9, alpha=0.9, seed=37,
              strategy="ori", protos=synth.warm_prototypes(7, seed=37), steps=1, tweak=tweak_nohard, block=3)

    # degenerate: a single class present -> loss 0 with dense zero grad (loss.py:116-117); prototypes still update
    def tweak_single(d, s):
        d["mask"][d["cls"].unsqueeze(1) != 1] = 0
    loss_case("loss_single_c4", B2=1, C=4, h=8, w=8, Q=4, Nn=8, temp=0.5, strong=0.97, alpha=0.99, seed=41,
              strategy="ori", protos=zeros(4), steps=1, tweak=tweak_single)
    # CityScapes class count
    loss_case("loss_cross_c19", B2=2, C=19, h=12, w=12, Q=16, Nn=48, temp=0.5, strong=0.8, alpha=0.99, seed=53,
              strategy="cross", protos=synth.warm_prototypes(19, seed=53), steps=1, block=3)

    print("stage 1/2 cases")
    stage12_case("stage12_mix_c21", kind="mix", B=2, C=21, h=9, w=9, H=33, W=33, temp=0.5, seed=61, zero_rows=(5,))
    stage12_case("stage12_cross_c19", kind="cross", B=1, C=19, h=7, w=10, H=25, W=37, temp=0.5, seed=67)
    stage12_case("stage12_ori_c21", kind="ori", B=2, C=21, h=9, w=9, H=33, W=33, temp=0.5, seed=71)

    print("attention-threshold loss cases")
    atl_case("atl_c21", B=3, C=21, H=23, W=19, thr=0.7, seed=91)
    atl_case("atl_c19", B=2, C=19, H=16, W=33, thr=0.97, seed=93)

    print("glue cases")
    glue_case("glue_mix", strategy="mix", B=2, C=21, H=33, W=33, h=9, w=9, weak=0.7, seed=81)
    glue_case("glue_cross", strategy="cross", B=2, C=19, H=25, W=37, h=7, w=10, weak=0.7, seed=83)
    glue_case("glue_ori", strategy="ori", B=1, C=21, H=40, W=40, h=10, w=10, weak=0.7, seed=87)

    print("augmentation hand-off cases")
    aug_case("aug_t2_scale", maps=2, B=3, C=21, H=41, W=47, crop=(32, 36), scale=(0.5, 1.5), augmentation=False, seed=7)
    aug_case("aug_t2_flip", maps=2, B=3, C=21, H=41, W=47, crop=(41, 47), scale=(1.0, 1.0), augmentation=True, seed=7,
             int_labels=True)
    aug_case("aug_t3_city", maps=3, B=3, C=19, H=45, W=52, crop=(50, 40), scale=(0.5, 2.0), augmentation=True, seed=9)
    aug_case("aug_t1_ori", maps=1, B=2, C=21, H=33, W=33, crop=(33, 33), scale=(0.5, 1.5), augmentation=True, seed=13)
    cut_case("cut_cutmix_2", maps=2, mode="cutmix", B=3, C=21, H=33, W=37, seed=21)
    cut_case("cut_cutmix_3", maps=3, mode="cutmix", B=2, C=19, H=30, W=30, seed=22)
    cut_case("cut_cutout_1", maps=1, mode="cutout", B=3, C=21, H=33, W=37, seed=23)
    cut_case("cut_classmix_2", maps=2, mode="classmix", B=3, C=21, H=33, W=37, seed=24)
    cut_world2_case("cut_cutmix_2_world2")
    cut_world2_case("cut_classmix_2_world2", mode="classmix", port=29643)


if __name__ == "__main__":
    main()
