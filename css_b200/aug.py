"""GPU-resident hand-off of the pseudo-label / confidence maps through the augmentation.

The reference sends every map of every image GPU -> PIL -> GPU twice per step (dataset_helpers/VOC.py:284-302 into
transform_* :64-274 and back, driven by batch_transform_* :312-352) and mixes images with a Python loop of small
kernels behind four all_gathers (generate_cut_gather_* :354-477).  For the maps, that trip is index arithmetic plus
an 8-bit quantisation, so here

  * the IMAGE keeps the reference's PIL path (bilinear resize, reflect pad, colour jitter, blur: BASELINE north_star
    leaves the augmentation of images on the PyTorch/PIL path) -- it is also what draws the random geometry, so the
    Python / torch / NumPy RNG streams are consumed exactly as the reference consumes them;
  * the MAPS never leave the GPU: `css_aug_index` + `css_aug_maps` replay the drawn geometry (Pillow NEAREST resize
    table, bottom/right padding with 255 / 0, crop, flip) and the 8-bit round trip (label -1 <-> 255, conf ->
    floor(conf * 255) / 255) bit-exactly;
  * CutMix / CutOut / ClassMix is one fused launch (`css_cut_mix`); with several ranks only rank 0's batch is
    broadcast, because the reference's partner index `(i + 1) % batch_size` always lands in rank 0's slice of the
    gathered batch (VOC.py:386,428,469).

`batch_transform`, `batch_transform_2`, `batch_transform_3`, `generate_cut_gather`, `generate_cut_gather_2`,
`generate_cut_gather_3` keep the reference's names, arguments and return values; `css_b200.install.install(...,
gpu_aug=True)` routes the model shells through them.
"""
import random

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import check, ptr, stream_ptr

_MEAN = (0.485, 0.456, 0.406)
_STD = (0.229, 0.224, 0.225)


# ---------------------------------------------------------------------------------------------------------------
# image side (CPU, PIL) -- also the source of the random geometry
# ---------------------------------------------------------------------------------------------------------------
def _tv():
    import torchvision.transforms as T
    import torchvision.transforms.functional as TF
    return T, TF


def _unnormalise(images):
    """ImageNet de-normalisation with the reference's two-step arithmetic (VOC.py:304-308), batched."""
    _, TF = _tv()
    x = TF.normalize(images, mean=[0., 0., 0.], std=[1 / s for s in _STD])
    return TF.normalize(x, mean=[-m for m in _MEAN], std=[1., 1., 1.])


def _augment_image(pil, crop_size, scale_size, augmentation):
    """One image through the reference's resize / pad / crop / jitter / blur / flip chain (the image lines of
    VOC.py:126-196).  Returns (normalised tensor [3,ch,cw], (resized_h, resized_w, top, left, flip)): the geometry is
    what the maps have to replay."""
    from PIL import ImageFilter
    T, TF = _tv()
    raw_w, raw_h = pil.size
    ratio = random.uniform(scale_size[0], scale_size[1])
    rh, rw = int(raw_h * ratio), int(raw_w * ratio)
    pil = TF.resize(pil, (rh, rw), T.InterpolationMode.BILINEAR)
    if crop_size == -1:
        crop_size = (raw_w, raw_h)
    ch, cw = int(crop_size[0]), int(crop_size[1])
    if ch > rh or cw > rw:
        pil = TF.pad(pil, padding=(0, 0, max(cw - rw, 0), max(ch - rh, 0)), padding_mode='reflect')
    top, left, _, _ = T.RandomCrop.get_params(pil, output_size=(ch, cw))
    pil = TF.crop(pil, top, left, ch, cw)
    flip = False
    if augmentation:
        if torch.rand(1) > 0.2:
            pil = T.ColorJitter((0.75, 1.25), (0.75, 1.25), (0.75, 1.25), (-0.25, 0.25))(pil)
        if torch.rand(1) > 0.5:
            pil = pil.filter(ImageFilter.GaussianBlur(radius=random.uniform(0.15, 1.15)))
        if torch.rand(1) > 0.5:
            pil = TF.hflip(pil)
            flip = True
    img = TF.normalize(TF.to_tensor(pil), mean=list(_MEAN), std=list(_STD))
    return img, (rh, rw, int(top), int(left), int(flip))


# ---------------------------------------------------------------------------------------------------------------
# map side (GPU)
# ---------------------------------------------------------------------------------------------------------------
def _label_arg(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"css_b200: `{name}` must be a CUDA tensor (no CPU fallback)")
    if t.dtype == torch.float32:
        return t.contiguous(), _lib.LABEL_F32
    if t.dtype == torch.int64:
        return t.contiguous(), _lib.LABEL_I64
    raise RuntimeError(f"css_b200: `{name}` must be float32 or int64, got {t.dtype}")


def _conf_arg(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"css_b200: `{name}` must be a CUDA tensor (no CPU fallback)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"css_b200: `{name}` must be float32, got {t.dtype}")
    return t.contiguous()


def transform_maps(labels, confs, geometry, crop_hw, max_resized=None):
    """Replays `geometry` (int tensor/array [B,5]: resized_h, resized_w, top, left, flip) on up to two label maps and up to
    two confidence maps [B,H,W] that stay on the GPU.  Returns (labels int64 [B,ch,cw] with -1 = ignore, confs float32).
    A CUDA `geometry` together with `max_resized` (an upper bound of every resized_h / resized_w) keeps the call free of
    host synchronisation, e.g. for CUDA-graph capture."""
    labels, confs = list(labels), list(confs)
    if not labels and not confs:
        raise RuntimeError("css_b200: transform_maps needs at least one map")
    if len(labels) > 2 or len(confs) > 2:
        raise RuntimeError("css_b200: transform_maps takes at most two label maps and two confidence maps")
    first = (labels + confs)[0]
    dev = first.device
    B, H, W = first.shape
    ch, cw = int(crop_hw[0]), int(crop_hw[1])
    la = [_label_arg(t, "label") for t in labels]
    if len(la) == 2 and la[0][1] != la[1][1]:
        raise RuntimeError("css_b200: the two label maps must have the same dtype")
    cf = [_conf_arg(t, "logits") for t in confs]
    for t in [x[0] for x in la] + cf:
        if tuple(t.shape) != (B, H, W):
            raise RuntimeError("css_b200: every map must be [B,H,W] with the same shape")
    if isinstance(geometry, torch.Tensor) and geometry.is_cuda:
        geo = geometry.to(torch.int32).contiguous()
        max_r = int(max_resized) if max_resized is not None else int(geo[:, :2].max().item())
    else:
        geo_host = np.ascontiguousarray(np.asarray(geometry, dtype=np.int32).reshape(B, 5))
        geo = torch.from_numpy(geo_host).pin_memory().to(dev, non_blocking=True)
        max_r = int(max(geo_host[:, 0].max(), geo_host[:, 1].max()))
    ymap = torch.empty((B, max_r), device=dev, dtype=torch.int32)
    xmap = torch.empty((B, max_r), device=dev, dtype=torch.int32)
    out_l = [torch.empty((B, ch, cw), device=dev, dtype=torch.int64) for _ in la]
    out_c = [torch.empty((B, ch, cw), device=dev, dtype=torch.float32) for _ in cf]
    lib = _lib.load()

    def at(seq, i):
        return ptr(seq[i]) if i < len(seq) else None

    with torch.cuda.device(dev):
        check(lib.css_aug_index(ptr(geo), B, H, W, max_r, ptr(ymap), ptr(xmap), stream_ptr()), "css_aug_index")
        check(lib.css_aug_maps(at([x[0] for x in la], 0), at([x[0] for x in la], 1), la[0][1] if la else _lib.LABEL_F32,
                               at(cf, 0), at(cf, 1), ptr(geo), ptr(ymap), ptr(xmap), B, H, W, max_r, ch, cw,
                               at(out_l, 0), at(out_l, 1), at(out_c, 0), at(out_c, 1), stream_ptr()), "css_aug_maps")
    return out_l, out_c


def _batch_transform(images, labels, confs, crop_size, scale_size, augmentation):
    if not images.is_cuda:
        raise RuntimeError("css_b200: `images` must be a CUDA tensor (no CPU fallback)")
    _, TF = _tv()
    dev = images.device
    host = _unnormalise(images).cpu()                       # ONE device->host copy for the whole batch
    out_img, geometry = [], []
    for k in range(host.shape[0]):
        img, geo = _augment_image(TF.to_pil_image(host[k]), crop_size, scale_size, augmentation)
        out_img.append(img)
        geometry.append(geo)
    crop_hw = out_img[0].shape[1:]
    out_l, out_c = transform_maps(labels, confs, np.asarray(geometry, np.int32), crop_hw)
    image_trans = torch.stack(out_img).pin_memory().to(dev, non_blocking=True)
    return image_trans, out_l, out_c


def batch_transform(images, labels, logits=None, crop_size=(512, 512), scale_size=(0.8, 1.0), augmentation=True):
    """Drop-in for dataset_helpers/VOC.py:312-323."""
    img, l, c = _batch_transform(images, [labels], [logits], crop_size, scale_size, augmentation)
    return img, l[0], c[0]


def batch_transform_2(images, labels, logits_1=None, logits_2=None, crop_size=(512, 512), scale_size=(0.8, 1.0),
                      augmentation=True):
    """Drop-in for dataset_helpers/VOC.py:325-337."""
    img, l, c = _batch_transform(images, [labels], [logits_1, logits_2], crop_size, scale_size, augmentation)
    return img, l[0], c[0], c[1]


def batch_transform_3(images, labels1, labels2, logits_1=None, logits_2=None, crop_size=(512, 512), scale_size=(0.8, 1.0),
                      augmentation=True):
    """Drop-in for dataset_helpers/VOC.py:339-352."""
    img, l, c = _batch_transform(images, [labels1, labels2], [logits_1, logits_2], crop_size, scale_size, augmentation)
    return img, l[0], l[1], c[0], c[1]


# ---------------------------------------------------------------------------------------------------------------
# CutOut / CutMix / ClassMix
# ---------------------------------------------------------------------------------------------------------------
CUT_MODES = {"cutout": 0, "cutmix": 1, "classmix": 2}


def draw_cut_box(image_h, image_w, ratio=2):
    """The region generate_cutout_mask zeroes (VOC.py:518-535), as (y0, y1, x0, x1) clipped to the image; consumes
    np.random exactly like the reference (three randint draws)."""
    area = image_h * image_w / ratio
    bw = np.random.randint(image_w / ratio + 1, image_w)
    bh = np.round(area / bw)
    x0 = np.random.randint(0, image_w - bw + 1)
    y0 = np.random.randint(0, image_h - bh + 1)
    x1, y1 = int(x0 + bw), int(y0 + bh)
    return int(y0), min(y1, image_h), int(x0), min(x1, image_w)


def draw_class_set(label_map):
    """ClassMix: half of the label values present in `label_map`, picked with torch.randperm (VOC.py:511-516)."""
    present = torch.unique(label_map)
    chosen = present[torch.randperm(len(present))][:len(present) // 2]
    return [int(v) for v in chosen.tolist()]


def _class_bits(values):
    """Label value v in [-1, 62] -> bit v + 1 of a 64-bit set."""
    bits = 0
    for v in values:
        if not -1 <= v <= 62:
            raise RuntimeError(f"css_b200: classmix label value {v} outside [-1, 62]")
        bits |= 1 << (v + 1)
    return bits - (1 << 64) if bits >= (1 << 63) else bits


def cut_mix(image, labels, confs, mode, boxes=None, class_sets=None, partner=None):
    """One fused launch: out[i] = keep ? own[i] : partner[(i+1) % B]  (cutout: 0 / -1 instead of the partner).
    boxes: int [B,4] (y0, y1, x0, x1) of the replaced region; class_sets: per image the label values that keep the own
    pixels; partner: (image, labels, confs) the partners come from (default: the same batch)."""
    if mode not in CUT_MODES:
        raise ValueError('mode must be in cutout, cutmix, or classmix')
    if not image.is_cuda:
        raise RuntimeError("css_b200: `image` must be a CUDA tensor (no CPU fallback)")
    dev = image.device
    image = _conf_arg(image, "image")
    B, CH, H, W = image.shape
    labels = [_label_arg(t, "label")[0] for t in labels]
    for t in labels:
        if t.dtype != torch.int64:
            raise RuntimeError("css_b200: cut_mix labels must be int64")
    confs = [_conf_arg(t, "logits") for t in confs]
    if len(labels) not in (1, 2) or len(confs) not in (1, 2):
        raise RuntimeError("css_b200: cut_mix takes one or two label maps and one or two confidence maps")
    p_image, p_labels, p_confs = (image, labels, confs) if partner is None else partner
    p_image = _conf_arg(p_image, "partner image")
    p_labels = [_label_arg(t, "partner label")[0] for t in p_labels]
    p_confs = [_conf_arg(t, "partner logits") for t in p_confs]
    if mode == "classmix":
        spec = torch.tensor([_class_bits(s) for s in class_sets], dtype=torch.int64).pin_memory().to(dev, non_blocking=True)
        box_t = None
    else:
        box_t = torch.from_numpy(np.ascontiguousarray(np.asarray(boxes, np.int32).reshape(B, 4))).pin_memory().to(dev, non_blocking=True)
        spec = None
    o_img = torch.empty_like(image)
    o_lab = [torch.empty_like(t) for t in labels]
    o_conf = [torch.empty_like(t) for t in confs]
    lib = _lib.load()

    def at(seq, i):
        return ptr(seq[i]) if i < len(seq) else None

    with torch.cuda.device(dev):
        check(lib.css_cut_mix(ptr(image), at(labels, 0), at(labels, 1), at(confs, 0), at(confs, 1),
                              ptr(p_image), at(p_labels, 0), at(p_labels, 1), at(p_confs, 0), at(p_confs, 1),
                              ptr(box_t), ptr(spec), CUT_MODES[mode], B, CH, H, W,
                              ptr(o_img), at(o_lab, 0), at(o_lab, 1), at(o_conf, 0), at(o_conf, 1), stream_ptr()),
              "css_cut_mix")
    return o_img, o_lab, o_conf


def _generate_cut_gather(image, labels, confs, mode):
    batch_size, _, H, W = image.shape
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    labels = [t.long() for t in labels]
    if mode == 'none':
        return image, labels, confs
    if mode not in CUT_MODES:
        raise ValueError('mode must be in cutout, cutmix, or classmix')
    if mode == 'cutout' and len(labels) == 2:
        # generate_cut_gather_3 never fills its second label list in this mode and dies in torch.cat (VOC.py:453-461)
        raise RuntimeError("css_b200: mode 'cutout' with two label maps is broken in the reference (VOC.py:453-461)")
    partner = None
    if world > 1 and mode != 'cutout':
        # partner (i + 1) % batch_size indexes the GATHERED batch, i.e. always rank 0's images (VOC.py:428)
        partner = tuple(_broadcast0(x) for x in (image, labels, confs))
    total = batch_size * world
    boxes = class_sets = None
    if mode == 'classmix':
        # every rank draws torch.randperm(number of label values present) for EVERY gathered image, in gathered order
        # (VOC.py:381,423,465 -> :511-516).  What the draws for other ranks' images need from those ranks is only that number (it
        # fixes how much of the CPU generator's stream a permutation consumes), so one all_gather of `batch_size` counts replaces
        # the gather of the label maps; this rank's own images get their real draws at their place in the sequence.
        own = range(rank * batch_size, (rank + 1) * batch_size)
        if world > 1:
            n_all = [None] * world                                # a few ints per rank: backend-agnostic object gather
            dist.all_gather_object(n_all, [int(torch.unique(labels[0][i]).numel()) for i in range(batch_size)])
            counts = [n for per_rank in n_all for n in per_rank]
        class_sets = []
        for i in range(total):
            if i in own:
                class_sets.append(draw_class_set(labels[0][i - rank * batch_size]))
            else:
                torch.randperm(int(counts[i]))                    # advance the generator exactly as the reference does
    else:
        ratio = 2
        drawn = [draw_cut_box(H, W, ratio) for _ in range(total)]      # every rank draws for the whole gathered batch
        boxes = drawn[rank * batch_size:(rank + 1) * batch_size]
    return cut_mix(image, labels, confs, mode, boxes=boxes, class_sets=class_sets, partner=partner)


def _broadcast0(x):
    if isinstance(x, (list, tuple)):
        return [_broadcast0(t) for t in x]
    buf = x.contiguous().clone()
    dist.broadcast(buf, src=0)
    return buf


def generate_cut_gather(image, label, logits, mode='cutout'):
    """Drop-in for dataset_helpers/VOC.py:354-391."""
    img, l, c = _generate_cut_gather(image, [label], [logits], mode)
    return img, l[0], c[0]


def generate_cut_gather_2(image, label, logits1, logits2, mode='cutout'):
    """Drop-in for dataset_helpers/VOC.py:393-434."""
    img, l, c = _generate_cut_gather(image, [label], [logits1, logits2], mode)
    return img, l[0], c[0], c[1]


def generate_cut_gather_3(image, label1, label2, logits1, logits2, mode='cutout'):
    """Drop-in for dataset_helpers/VOC.py:436-477."""
    img, l, c = _generate_cut_gather(image, [label1, label2], [logits1, logits2], mode)
    return img, l[0], l[1], c[0], c[1]
