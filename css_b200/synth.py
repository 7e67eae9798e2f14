"""Synthetic inputs of VOC / CityScapes shape for the contrastive path.

No datasets or pretrained weights exist offline, so every test, golden bundle and benchmark draws its
inputs here: a blocky class map (all classes present), class-informative 256-d representations,
noisy logits, a Bernoulli valid mask.  CPU torch generators only, so that the same seed gives the
same tensors in the build container and on the GPU box.
"""
import math

import torch
import torch.nn.functional as F


def _gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def class_map(B, C, h, w, g, ignore_frac=0.05, block=8):
    """Blocky segmentation map [B,h,w] int64 in {-1,0..C-1}: coarse randint, nearest up-sampling, 5 % ignore."""
    ch, cw = math.ceil(h / block), math.ceil(w / block)
    coarse = torch.randint(0, C, (B, 1, ch, cw), generator=g).float()
    cls = F.interpolate(coarse, size=(h, w), mode="nearest")[:, 0].long()
    if ignore_frac > 0:
        cls = torch.where(torch.rand(B, h, w, generator=g) < ignore_frac, torch.full_like(cls, -1), cls)
    return cls


def rep_for(cls, C, D, g, centers=None, signal=0.5):
    """rep [B,D,h,w] fp32 = signal * G[class] + N(0,1); ignored pixels are pure noise."""
    B, h, w = cls.shape
    if centers is None:
        centers = torch.randn(C, D, generator=g)
    idx = cls.clamp_min(0)
    mean = centers[idx] * (cls >= 0).unsqueeze(-1)                # [B,h,w,D]
    rep = signal * mean + torch.randn(B, h, w, D, generator=g)
    return rep.permute(0, 3, 1, 2).contiguous(), centers


def logits_for(cls, C, g, flip_frac=0.3):
    """logits [B,C,h,w] = 3*onehot(class') + 1.5*N(0,1), class' = class re-drawn on 30 % of the pixels."""
    B, h, w = cls.shape
    noisy = torch.where(torch.rand(B, h, w, generator=g) < flip_frac,
                        torch.randint(0, C, (B, h, w), generator=g), cls.clamp_min(0))
    onehot = F.one_hot(noisy, C).permute(0, 3, 1, 2).float()
    return (3.0 * onehot + 1.5 * torch.randn(B, C, h, w, generator=g)).contiguous()


def onehot_label(cls, C, zero_ignored=False):
    """Dense one-hot [B,C,h,w] fp32.  zero_ignored=False follows utils.py:116-125 (-1 -> class 0, relies on the
    mask); True follows utils.py:127-136 + mix_label.py:181-182 (ignored pixels get an all-zero row)."""
    if zero_ignored:
        return F.one_hot(cls + 1, C + 1).permute(0, 3, 1, 2)[:, 1:].float().contiguous()
    return F.one_hot(cls.clamp_min(0), C).permute(0, 3, 1, 2).float().contiguous()


def student_batch(B2, C, h, w, seed=3407, D=256, strategy="ori", mask_keep=0.8, block=8):
    """Inputs of Contrast_Loss.forward: rep_all, label_all, mask_all, prob_all (+ the class map).

    strategy 'ori': prob = softmax(logits) (ori_pseudo.py:180); 'mix'/'cross': prob = softmax(cos(rep, proto)/temp)
    is produced by the path itself (proto_softmax_sim), so prob here is only the logits-softmax placeholder."""
    g = _gen(seed)
    cls = class_map(B2, C, h, w, g, block=block)
    rep, centers = rep_for(cls, C, D, g)
    logits = logits_for(cls, C, g)
    mask = ((torch.rand(B2, 1, h, w, generator=g) < mask_keep) & (cls >= 0).unsqueeze(1)).float()
    label = onehot_label(cls, C, zero_ignored=(strategy == "mix"))
    prob = torch.softmax(logits, dim=1)
    return dict(rep=rep, label=label, mask=mask, prob=prob, logits=logits, cls=cls, centers=centers)


def teacher_batch(B, C, h, w, seed=3407, D=256, block=8):
    """Inputs of stage 1/2: the EMA teacher's rep_u [B,D,h,w] and pred_u [B,C,h,w] on the unlabelled images."""
    g = _gen(seed + 7919)
    cls = class_map(B, C, h, w, g, ignore_frac=0.0, block=block)
    rep_u, centers = rep_for(cls, C, D, g)
    pred_u = logits_for(cls, C, g)
    return dict(rep_u=rep_u, pred_u=pred_u, cls=cls, centers=centers)


def warm_prototypes(C, D=256, seed=3407, zero_rows=()):
    """Non-zero prototypes (as after a few EMA steps); rows in zero_rows stay zero (first-touch branch)."""
    g = _gen(seed + 104729)
    p = 0.5 * torch.randn(C, D, generator=g)
    for r in zero_rows:
        p[r] = 0
    return p
