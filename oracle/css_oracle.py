"""CPU restatement (numpy, fp32) of the CSS representation-space hot path.

TEST INFRASTRUCTURE -- this is the parity oracle, not a product path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  Nothing under ``css_b200/`` imports ``oracle``; the product path
raises if the CUDA library is missing.

Parity status: PINNED by execution of the live reference.  The reference ships no golden
vectors or tests of its own, so ``oracle/gen_golden.py`` runs the unmodified
reference (``/root/reference``, WangChangqi98/CSS) on CPU in the build container, records its
inputs, RNG draws, outputs and autograd gradients into ``tests/golden/*.npz`` and
``tests/test_oracle_golden.py`` checks every function below against those bundles.

Every function cites the reference lines it restates (paths relative to the reference root).
Third-party arithmetic the reference relies on (none of it vendored): torch ``F.normalize``,
``torch.mm``, ``F.interpolate(bilinear, align_corners=True)``, ``softmax``, ``max``,
``cosine_similarity`` (per-norm eps clamp, torch>=1.12 semantics; the reference pins
torch==1.7.1 which cannot run on sm_100), ``F.cross_entropy``, ``torch.randint``,
``torch.multinomial`` and ``numpy.random.randint``.  Their published semantics are restated
here in numpy fp32.
"""
import numpy as np

f32 = np.float32


# --------------------------------------------------------------------------------------
# stage 1 / 1b : cosine + softmax similarity against the class prototypes
# --------------------------------------------------------------------------------------
def l2_normalize(x, axis=-1, eps=1e-12):
    """torch.nn.functional.normalize: x / max(||x||_2, eps).
    generalframeworks/networks/ddp_model.py:106-107 (and :150, :191-192, :233)."""
    x = np.asarray(x, dtype=f32)
    n = np.sqrt(np.sum(x * x, axis=axis, keepdims=True, dtype=f32)).astype(f32)
    return (x / np.maximum(n, f32(eps))).astype(f32)


def cos_sim_map(rep, prototypes):
    """Cosine similarity of every pixel's D-vector against every class prototype.
    ddp_model.py:104-110 (Model_mix), :189-195 (Model_cross), :147-153 / :230-236 (student).
    rep [B,D,h,w] f32, prototypes [C,D] f32 -> sim [B,C,h,w] f32.  Zero prototype rows give 0."""
    rep = np.asarray(rep, dtype=f32)
    B, D, h, w = rep.shape
    x = l2_normalize(rep.transpose(0, 2, 3, 1), axis=-1)          # :105-106
    p = l2_normalize(np.asarray(prototypes, dtype=f32), axis=-1)  # :107
    sim = x.reshape(B * h * w, D) @ p.T                           # :108-109
    return np.ascontiguousarray(sim.reshape(B, h, w, -1).transpose(0, 3, 1, 2)).astype(f32)  # :110


def bilinear_upsample(x, out_hw):
    """F.interpolate(x, size=out_hw, mode='bilinear', align_corners=True) (ddp_model.py:111,113;
    :36 in Model_ori_pseudo; :196,198 in Model_cross).  Restates ATen's upsample_bilinear2d:
    scale = (in-1)/(out-1) in fp32, src = scale*dst, i0 = int(src), i1 = i0 + (i0 < in-1),
    lam = src - i0, out = (1-ly)*((1-lx)*v00 + lx*v01) + ly*((1-lx)*v10 + lx*v11)."""
    x = np.asarray(x, dtype=f32)
    B, C, h, w = x.shape
    H, W = out_hw

    def axis(inp, out):
        scale = f32(inp - 1) / f32(out - 1) if out > 1 else f32(0)
        src = (scale * np.arange(out, dtype=f32)).astype(f32)
        i0 = src.astype(np.int64)
        i1 = i0 + (i0 < inp - 1)
        lam1 = (src - i0.astype(f32)).astype(f32)
        lam0 = (f32(1) - lam1).astype(f32)
        return i0, i1, lam0, lam1

    y0, y1, hy, ly = axis(h, H)
    x0, x1, hx, lx = axis(w, W)
    hx = hx[None, None, None, :]
    lx = lx[None, None, None, :]
    hy = hy[None, None, :, None]
    ly = ly[None, None, :, None]
    top = x[:, :, y0, :]
    bot = x[:, :, y1, :]
    t = (hx * top[:, :, :, x0]).astype(f32) + (lx * top[:, :, :, x1]).astype(f32)
    b = (hx * bot[:, :, :, x0]).astype(f32) + (lx * bot[:, :, :, x1]).astype(f32)
    return ((hy * t).astype(f32) + (ly * b).astype(f32)).astype(f32)


def softmax(x, axis=1):
    """torch.softmax in fp32: exp(x - max) / sum."""
    x = np.asarray(x, dtype=f32)
    m = np.max(x, axis=axis, keepdims=True)
    e = np.exp((x - m).astype(f32)).astype(f32)
    s = np.sum(e, axis=axis, keepdims=True, dtype=f32).astype(f32)
    return (e / s).astype(f32)


def softmax_max(x, temp=None):
    """torch.max(softmax(x / temp, dim=1), dim=1) -> (conf f32, label i64); first index on ties.
    ddp_model.py:112 (temp) and :114 (no temp)."""
    x = np.asarray(x, dtype=f32)
    if temp is not None:
        x = (x / f32(temp)).astype(f32)
    p = softmax(x, axis=1)
    return np.max(p, axis=1).astype(f32), np.argmax(p, axis=1).astype(np.int64)


def rep_pseudo_label(rep_u, prototypes, temp, out_hw):
    """Representation-space pseudo label + confidence at crop resolution. ddp_model.py:104-112."""
    sim = cos_sim_map(rep_u, prototypes)
    return softmax_max(bilinear_upsample(sim, out_hw), temp) + (sim,)


def cls_pseudo_label(pred_u, out_hw):
    """Logit-space pseudo label + confidence at crop resolution. ddp_model.py:113-114 (:36-37, :198-199)."""
    return softmax_max(bilinear_upsample(pred_u, out_hw), None)


def mix_fuse(label_cls, label_rep, num_classes):
    """Collaborative (mix-label) fusion, ddp_model.py:115-118: keep the logit-space label where the two
    spaces agree, otherwise 255.  Output is float32 as in the reference (int64 - float -> float)."""
    label_cls = np.asarray(label_cls, dtype=np.int64)
    disagree = (label_cls != np.asarray(label_rep, dtype=np.int64)).astype(f32)      # :115-116
    out = (label_cls.astype(f32) - disagree * f32(num_classes)).astype(f32)          # :117
    out[out < 0] = f32(255)                                                          # :118
    return out


def proto_softmax_sim(rep_all, prototypes, temp):
    """Student softmax-similarity map ("tilt" indicator) at rep resolution. ddp_model.py:147-154."""
    sim = cos_sim_map(rep_all, prototypes)
    return softmax((sim / f32(temp)).astype(f32), axis=1)


# --------------------------------------------------------------------------------------
# stage 2 glue : thresholds, one-hot, nearest down-sampling (script train() bodies)
# --------------------------------------------------------------------------------------
def label_onehot(inputs, num_class):
    """generalframeworks/utils.py:116-125: relu then scatter -> [B,C,H,W] f32 (-1 lands in class 0)."""
    idx = np.maximum(np.asarray(inputs, dtype=np.int64), 0)
    B, H, W = idx.shape
    out = np.zeros((B, num_class, H, W), dtype=f32)
    np.put_along_axis(out, idx[:, None], f32(1), axis=1)
    return out


def label_onehot_2(inputs, num_class):
    """generalframeworks/utils.py:127-136: shift by +1, C+1 channels (caller drops channel 0)."""
    idx = np.asarray(inputs, dtype=np.int64) + 1
    B, H, W = idx.shape
    out = np.zeros((B, num_class + 1, H, W), dtype=f32)
    np.put_along_axis(out, idx[:, None], f32(1), axis=1)
    return out


def nearest_downsample(x, out_hw):
    """F.interpolate(x, size=out_hw, mode='nearest'): src = min(floor(dst * (in/out)), in-1), fp32 scale."""
    x = np.asarray(x)
    H, W = x.shape[-2:]
    h, w = out_hw

    def idx(inp, out):
        scale = f32(inp) / f32(out)
        return np.minimum(np.floor(np.arange(out, dtype=f32) * scale).astype(np.int64), inp - 1)

    return x[..., idx(H, h), :][..., idx(W, w)]


def threshold_glue(train_l_label, train_u_aug_label, train_u_aug_logits_cls, weak_threshold, num_class,
                   out_hw, strategy):
    """mix_label.py:175-183 (strategy 'mix'), cross_label.py:178-185 ('cross'), ori_pseudo.py:171-178 ('ori').
    Returns (label_all [2B,C,h,w] f32, mask_all [2B,1,h,w] f32)."""
    train_l_label = np.asarray(train_l_label, dtype=np.int64)
    train_u_aug_label = np.asarray(train_u_aug_label, dtype=np.int64)
    u_mask = (np.asarray(train_u_aug_logits_cls, dtype=f32) >= f32(weak_threshold)).astype(f32)
    mask_all = np.concatenate([(train_l_label[:, None] >= 0).astype(f32), u_mask[:, None]])
    mask_all = nearest_downsample(mask_all, out_hw)
    label_l = nearest_downsample(label_onehot(train_l_label, num_class), out_hw)
    if strategy == "mix":
        label_u = nearest_downsample(label_onehot_2(train_u_aug_label, num_class), out_hw)[:, 1:]
    else:
        label_u = nearest_downsample(label_onehot(train_u_aug_label, num_class), out_hw)
    return np.concatenate([label_l, label_u]).astype(f32), mask_all.astype(f32)


# --------------------------------------------------------------------------------------
# stage 3 + 4 : Contrast_Loss (prototype EMA, selection, sampling, scoring, CE, backward)
# --------------------------------------------------------------------------------------
def negative_index_sampler(samp_num, seg_num_list, randint=None):
    """generalframeworks/loss/loss.py:410-418, same loop and same numpy.random.randint call order, so that
    with an identically seeded global RandomState it reproduces the reference's indices."""
    randint = np.random.randint if randint is None else randint
    negative_index = []
    for i in range(samp_num.shape[0]):
        for j in range(samp_num.shape[1]):
            negative_index += randint(low=sum(seg_num_list[:j]), high=sum(seg_num_list[:j + 1]),
                                      size=int(samp_num[i, j])).tolist()
    return negative_index


def cosine_similarity(a, b, axis, eps=1e-8):
    """torch.cosine_similarity (>=1.12): sum((a/max(|a|,eps)) * (b/max(|b|,eps)))."""
    a = np.asarray(a, dtype=f32)
    b = np.asarray(b, dtype=f32)
    na = np.maximum(np.sqrt(np.sum(a * a, axis=axis, keepdims=True, dtype=f32)), f32(eps)).astype(f32)
    nb = np.maximum(np.sqrt(np.sum(b * b, axis=axis, keepdims=True, dtype=f32)), f32(eps)).astype(f32)
    return np.sum((a / na) * (b / nb), axis=axis, dtype=f32).astype(f32)


def proto_class_prob(proto_rep, k, temp):
    """loss.py:133-135: softmax(cos(proto_k, proto_others)/temp) with others in rotated order k+1..V-1,0..k-1."""
    V = proto_rep.shape[0]
    id_mask = np.concatenate([np.arange(k, V), np.arange(0, k)])
    sim = cosine_similarity(proto_rep[id_mask[0]][None], proto_rep[id_mask[1:]], axis=1)
    return softmax((sim / f32(temp)).astype(f32), axis=0)


class ReferenceOrderSampler:
    """Draws anchors / negative classes / negative indices with the same library calls, in the same order, as
    loss.py:127,136-140 (torch CPU generator, torch.multinomial via Categorical.sample, global numpy RandomState).
    Seed torch and numpy exactly like the reference run to reproduce its draws."""

    def anchors(self, n_hard, Q):
        import torch
        return torch.randint(n_hard, size=(Q,)).numpy().astype(np.int64)

    def negatives(self, proto_prob, Q, Nn, negative_num_list):
        import torch
        dist = torch.distributions.categorical.Categorical(probs=torch.from_numpy(np.asarray(proto_prob)))
        samp_class = dist.sample(sample_shape=[Q, Nn])
        samp_num = torch.stack([(samp_class == c).sum(1) for c in range(len(proto_prob))], dim=1)   # loss.py:138
        return np.asarray(negative_index_sampler(samp_num, negative_num_list), dtype=np.int64)


class RecordedDraws:
    """Feeds recorded draws (lists with one entry per scored class, in scoring order)."""

    def __init__(self, anchor_idx, neg_idx):
        self.anchor_idx = [np.asarray(a, dtype=np.int64) for a in anchor_idx]
        self.neg_idx = [np.asarray(a, dtype=np.int64).reshape(-1) for a in neg_idx]
        self._a = 0
        self._n = 0

    def anchors(self, n_hard, Q):
        out = self.anchor_idx[self._a]
        self._a += 1
        assert out.shape == (Q,) and (out.size == 0 or out.max() < n_hard)
        return out

    def negatives(self, proto_prob, Q, Nn, negative_num_list):
        out = self.neg_idx[self._n]
        self._n += 1
        assert out.shape == (Q * Nn,) and (out.size == 0 or out.max() < sum(negative_num_list))
        return out


def class_statistics(rep, label, mask):
    """Per-class feature sums [C,D] and valid-pixel counts [C] of one rank: the quantities whose all-reduce
    replaces the reference's all_gather (loss.py:77,81,102): mean_c = sum_ranks(sums_c) / sum_ranks(cnt_c)."""
    rep = np.asarray(rep, dtype=f32)
    valid = (np.asarray(label, dtype=f32) * np.asarray(mask, dtype=f32)) != 0          # [B2,C,h,w]
    x = rep.transpose(0, 2, 3, 1)
    C = valid.shape[1]
    sums = np.zeros((C, rep.shape[1]), dtype=f32)
    cnt = np.zeros((C,), dtype=f32)
    for c in range(C):
        rows = x[valid[:, c]]
        cnt[c] = rows.shape[0]
        if rows.shape[0]:
            sums[c] = rows.sum(axis=0, dtype=f32)
    return sums, cnt


def contrast_loss(rep, label, mask, prob, prototypes, *, num_queries, num_negatives, temp=0.5,
                  strong_threshold=0.97, alpha=0.99, sampler=None, rep_gather=None, valid_gather=None,
                  want_grad=True):
    """Contrast_Loss.forward (loss.py:75-149) and the closed-form gradient autograd gives for it.

    rep [B2,D,h,w], label [B2,C,h,w], mask [B2,1,h,w], prob [B2,C,h,w], prototypes [C,D] (UPDATED IN PLACE, :105,:108).
    rep_gather / valid_gather: what concat_all_gather returns on a multi-rank job (:77,:81); default = local.
    Returns (loss f32, grad_rep [B2,D,h,w] f32 or None, info dict with every intermediate the CUDA path exposes)."""
    sampler = ReferenceOrderSampler() if sampler is None else sampler
    rep = np.asarray(rep, dtype=f32)
    label = np.asarray(label, dtype=f32)
    mask = np.asarray(mask, dtype=f32)
    prob = np.asarray(prob, dtype=f32)
    assert prototypes.dtype == f32
    B2, D, h, w = rep.shape
    C = label.shape[1]
    Q, Nn = num_queries, num_negatives
    valid_all = (label * mask).astype(f32)                               # :80
    rep_prt = rep if rep_gather is None else np.asarray(rep_gather, dtype=f32)
    valid_prt = valid_all if valid_gather is None else np.asarray(valid_gather, dtype=f32)
    x = rep.transpose(0, 2, 3, 1).reshape(-1, D)                         # :85   pixel-major rows, id = (b*h + y)*w + x
    x_prt = rep_prt.transpose(0, 2, 3, 1).reshape(-1, D)                 # :86

    present, num_list, valid_ids, hard_ids, proto_rep_list = [], [], [], [], []
    for i in range(C):                                                   # :93
        valid = valid_all[:, i].reshape(-1)
        if valid.sum(dtype=f32) == 0:                                    # :96
            continue
        vb = valid != 0
        hard = (prob[:, i].reshape(-1) < f32(strong_threshold)) & vb     # :99
        mean = x_prt[valid_prt[:, i].reshape(-1) != 0].mean(axis=0, dtype=f32).astype(f32)   # :102
        if prototypes[i].sum(dtype=f32) == 0:                            # :103
            prototypes[i] = mean                                         # :105
        else:
            prototypes[i] = (f32(alpha) * prototypes[i] + f32(1 - alpha) * mean).astype(f32)  # :108
        proto_rep_list.append(prototypes[i].copy())                      # :104 / :109
        present.append(i)
        valid_ids.append(np.flatnonzero(vb))                             # :111 (row-major order)
        hard_ids.append(np.flatnonzero(hard))                            # :112
        num_list.append(int(vb.sum()))                                   # :113

    V = len(num_list)
    info = dict(present=present, num_list=num_list, valid_ids=valid_ids, hard_ids=hard_ids, V=V,
                anchor_pixels=[], class_loss=[], scored=[], anchor_idx=[], neg_idx=[], neg_pixels=[])
    grad = np.zeros_like(rep) if want_grad else None
    if V <= 1:                                                           # :116-117
        return f32(0), grad, info

    proto_rep = np.stack(proto_rep_list).astype(f32)                     # :120
    grad_rows = np.zeros((B2 * h * w, D), dtype=f32) if want_grad else None
    loss = f32(0)
    for k in range(V):                                                   # :124
        if len(hard_ids[k]) == 0:                                        # :125,:129-130
            continue
        sample_idx = sampler.anchors(len(hard_ids[k]), Q)                # :127
        anchor_px = hard_ids[k][sample_idx]
        a = x[anchor_px]                                                 # :128  [Q,D]
        proto_prob = proto_class_prob(proto_rep, k, temp)                # :133-135
        negative_num_list = num_list[k + 1:] + num_list[:k]              # :139
        neg_index = sampler.negatives(proto_prob, Q, Nn, negative_num_list)            # :136-140
        negcat = np.concatenate(valid_ids[k + 1:] + valid_ids[:k])       # :141
        neg_px = negcat[neg_index].reshape(Q, Nn)
        cand = np.concatenate([np.broadcast_to(proto_rep[k], (Q, 1, D)), x[neg_px]], axis=1)   # :142-144 [Q,1+Nn,D]
        na = np.maximum(np.sqrt(np.sum(a * a, axis=1, keepdims=True, dtype=f32)), f32(1e-8)).astype(f32)
        nc = np.maximum(np.sqrt(np.sum(cand * cand, axis=2, keepdims=True, dtype=f32)), f32(1e-8)).astype(f32)
        a_hat = (a / na).astype(f32)
        c_hat = (cand / nc).astype(f32)
        cos = np.einsum("qd,qjd->qj", a_hat, c_hat).astype(f32)           # :146
        z = (cos / f32(temp)).astype(f32)
        m = z.max(axis=1, keepdims=True)
        e = np.exp((z - m).astype(f32)).astype(f32)
        s = e.sum(axis=1, keepdims=True, dtype=f32)
        lse = (np.log(s) + m).astype(f32)
        cls_loss = np.mean((lse[:, 0] - z[:, 0]).astype(f32), dtype=f32)  # :147
        loss = f32(loss + cls_loss)
        info["scored"].append(k)
        info["anchor_idx"].append(sample_idx)
        info["neg_idx"].append(np.asarray(neg_index, dtype=np.int64))
        info["anchor_pixels"].append(anchor_px)
        info["neg_pixels"].append(neg_px)
        info["class_loss"].append(cls_loss)
        if want_grad:
            # d/da of mean_q CE(cos/temp, 0) / V  
            pi = (e / s).astype(f32)
            g = pi.copy()
            g[:, 0] -= f32(1)
            g = (g / f32(Q * V * temp)).astype(f32)
            ga = (np.einsum("qj,qjd->qd", g, c_hat) - (g * cos).sum(axis=1, keepdims=True) * a_hat) / na
            np.add.at(grad_rows, anchor_px, ga.astype(f32))
    loss = f32(loss / f32(V))                                            # :149
    if want_grad:
        grad = np.ascontiguousarray(grad_rows.reshape(B2, h, w, D).transpose(0, 3, 1, 2))
    info["proto_rep"] = proto_rep
    return loss, grad, info


# --------------------------------------------------------------------------------------
# SURVEY.md 8(f)-3 : Attention_Threshold_Loss (logit-space unsupervised loss)
# --------------------------------------------------------------------------------------
def attention_threshold_loss(pred, pseudo_label, logits, strong_threshold, want_grad=True):
    """generalframeworks/loss/loss.py:53-64 and the gradient autograd gives for it.
    pred [B,C,H,W] f32, pseudo_label [B,H,W] int (-1 = ignore), logits [B,H,W] f32 (confidences)."""
    pred = np.asarray(pred, dtype=f32)
    lab = np.asarray(pseudo_label, dtype=np.int64)
    conf = np.asarray(logits, dtype=f32)
    B, C, H, W = pred.shape
    valid = lab >= 0                                                                      # :55
    weighting = ((conf.reshape(B, -1) >= f32(strong_threshold)).sum(-1).astype(f32)
                 / valid.reshape(B, -1).sum(-1).astype(f32)).astype(f32)                  # :56
    m = pred.max(axis=1, keepdims=True)
    e = np.exp((pred - m).astype(f32)).astype(f32)
    ssum = e.sum(axis=1, keepdims=True, dtype=f32)
    lse = (m + np.log(ssum))[:, 0].astype(f32)
    picked = np.take_along_axis(pred, np.maximum(lab, 0)[:, None], axis=1)[:, 0]
    loss = np.where(valid, lse - picked, f32(0)).astype(f32)                              # :59 (ignore_index=-1 -> 0)
    sel = loss > 0                                                                        # :60
    out = np.mean((weighting[:, None, None] * loss)[sel], dtype=f32) if sel.any() else f32(np.nan)
    grad = None
    if want_grad:
        n = f32(sel.sum())
        soft = (e / ssum).astype(f32)
        onehot = np.zeros_like(pred)
        np.put_along_axis(onehot, np.maximum(lab, 0)[:, None], f32(1), axis=1)
        # unselected pixels get exactly 0 (also when their image's weighting is nan: nll_loss backward skips ignored targets)
        coef = np.where(sel, (weighting[:, None, None] / n).astype(f32), f32(0))
        grad = ((soft - onehot) * coef[:, None]).astype(f32)
    return f32(out), grad


# ---------------------------------------------------------------------------------------------------------------
# the label / confidence maps' trip through the augmentation (tensor -> PIL 'L' image -> nearest
# resize -> pad -> crop -> flip -> tensor), restated as index arithmetic.  dataset_helpers/VOC.py:126-196 (transform_2;
# transform :64-124 and transform_3 :198-274 are the same with other map counts), :284-302 (tensor_to_pil_*).
# The arithmetic is Pillow's (NEAREST resize = ImagingScaleAffine, Pillow 9.1 .. 12.x) and torchvision's
# (to_pil_image: mul(255).byte(); to_tensor: byte / 255), neither vendored in the reference; pinned by
# tests/golden/aug_*.npz, generated by running the reference's own batch_transform_* here.
# ---------------------------------------------------------------------------------------------------------------
def pil_nearest_table(n_in, n_out):
    """Source index of every output pixel of Image.resize((.., n_out), NEAREST) along one axis.
    Pillow walks `pos = 0.5 * a; pos += a` in double with a = n_in / n_out and truncates (Geometry.c, affine scale path),
    so the table is a running sum, not (x + 0.5) * a."""
    a = np.float64(n_in) / np.float64(n_out)
    pos = np.float64(0.0) + a * np.float64(0.5)
    out = np.empty(n_out, np.int64)
    for x in range(n_out):
        out[x] = int(pos)
        pos = pos + a
    return np.minimum(out, n_in - 1)


def label_to_byte(label):
    """tensor_to_pil_*: (label.float() / 255) -> to_pil_image -> mul(255).byte().  Identity on 0..255, -1 wraps to 255."""
    f = (np.asarray(label).astype(np.float32) / np.float32(255.0)).astype(np.float32)
    return (f * np.float32(255.0)).astype(np.float32).astype(np.int64).astype(np.uint8)


def conf_to_byte(conf):
    """to_pil_image of a float map: mul(255).byte() (truncation): the 8-bit quantisation of the confidences."""
    return (np.asarray(conf, np.float32) * np.float32(255.0)).astype(np.float32).astype(np.int64).astype(np.uint8)


def byte_to_label(b):
    """(to_tensor(label) * 255).long(), 255 -> -1  (VOC.py:184-185)."""
    v = ((b.astype(np.float32) / np.float32(255.0)).astype(np.float32) * np.float32(255.0)).astype(np.float32).astype(np.int64)
    v[v == 255] = -1
    return v


def byte_to_conf(b):
    return (b.astype(np.float32) / np.float32(255.0)).astype(np.float32)


def aug_maps(labels, confs, geometry, crop_hw):
    """labels / confs: lists of [B,H,W] maps.  geometry: int array [B,5] = (resized_h, resized_w, top, left, flip) per
    image, i.e. what transform_* drew.  Returns (labels int64 with -1 = ignore, confs float32 multiples of 1/255)."""
    ch, cw = crop_hw
    B, H, W = np.asarray((labels + confs)[0]).shape
    out_l = [np.empty((B, ch, cw), np.int64) for _ in labels]
    out_c = [np.empty((B, ch, cw), np.float32) for _ in confs]
    for b in range(B):
        rh, rw, top, left, flip = (int(v) for v in geometry[b])
        ymap, xmap = pil_nearest_table(H, rh), pil_nearest_table(W, rw)
        ph, pw = max(rh, ch), max(rw, cw)                         # bottom / right constant padding (VOC.py:143-151)
        for maps, outs, to_byte, from_byte, fill in ((labels, out_l, label_to_byte, byte_to_label, 255),
                                                    (confs, out_c, conf_to_byte, byte_to_conf, 0)):
            for m, o in zip(maps, outs):
                img = np.full((ph, pw), fill, np.uint8)
                img[:rh, :rw] = to_byte(np.asarray(m[b]))[ymap][:, xmap]
                img = img[top:top + ch, left:left + cw]
                if flip:
                    img = img[:, ::-1]
                o[b] = from_byte(img)
    return out_l, out_c


def cut_mix(image, labels, confs, mode, boxes=None, class_sets=None, partner=None):
    """generate_cut_gather_* on one rank's slice (VOC.py:354-477).  boxes [B,4] = (y0, y1, x0, x1) of the region whose
    mask is 0; class_sets: per image, the label values whose pixels KEEP the own image (classmix).  partner =
    (image, labels, confs) of the batch the (i+1) % B partner is taken from (rank 0's; the own batch when None)."""
    B, _, H, W = image.shape
    p_img, p_lab, p_conf = partner if partner is not None else (image, labels, confs)
    o_img = np.empty_like(image)
    o_lab = [np.empty((B, H, W), np.int64) for _ in labels]
    o_conf = [np.empty((B, H, W), np.float32) for _ in confs]
    for i in range(B):
        if mode == "classmix":
            keep = np.isin(labels[0][i], np.asarray(class_sets[i]))
        else:
            keep = np.ones((H, W), bool)
            y0, y1, x0, x1 = (int(v) for v in boxes[i])
            keep[y0:y1, x0:x1] = False
        if mode == "cutout":
            o_img[i] = np.where(keep, image[i], 0)
            o_lab[0][i] = np.where(keep, labels[0][i], -1)
            for k in range(1, len(labels)):
                o_lab[k][i] = labels[k][i]
            for k in range(len(confs)):
                o_conf[k][i] = np.where(keep, confs[k][i], 0)
            continue
        j = (i + 1) % B
        o_img[i] = np.where(keep, image[i], p_img[j])
        for k in range(len(labels)):
            o_lab[k][i] = np.where(keep, labels[k][i], p_lab[k][j])
        for k in range(len(confs)):
            o_conf[k][i] = np.where(keep, confs[k][i], p_conf[k][j])
    return o_img, o_lab, o_conf
