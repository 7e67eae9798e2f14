"""Model_ori_pseudo / Model_mix / Model_cross shells: same constructors, attributes and forward return tuples as
generalframeworks/networks/ddp_model.py:8-239, with the inline representation-space blocks (:104-118, :147-154,
:189-199, :230-237, :36-37) replaced by the CUDA ops of css_b200.ops.

The DeepLabv3+ network (cuDNN) and the PIL / CutMix augmentation stay on the reference's PyTorch path: the shells take
them from `css_b200.models.hooks`, which `css_b200.install.install()` fills from the reference's own modules
(tests fill it with stubs).
"""
import copy
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops

# filled by install() with the reference's DeepLabv3Plus_with_rep and VOC.py augmentation functions
hooks = types.SimpleNamespace(network_factory=None, batch_transform=None, batch_transform_2=None, batch_transform_3=None,
                              generate_cut_gather=None, generate_cut_gather_2=None, generate_cut_gather_3=None)


def _hook(name):
    fn = getattr(hooks, name)
    if fn is None:
        raise RuntimeError(f"css_b200.models.hooks.{name} is not set: call css_b200.install.install() (needs the reference "
                           "on sys.path) or assign the hook yourself")
    return fn


class _TeacherStudent(nn.Module):
    def __init__(self, base_encoder, num_classes, output_dim, ema_alpha, config):
        super().__init__()
        self.model = _hook("network_factory")(base_encoder, num_classes=num_classes, output_dim=output_dim, dilate_scale=8)
        self.num_classes = num_classes
        self.step = 0
        self.ema_model = copy.deepcopy(self.model)
        for p in self.ema_model.parameters():
            p.requires_grad = False
        self.alpha = ema_alpha
        print('EMA model has been prepared. Alpha = {}'.format(self.alpha))
        self.config = config

    @torch.no_grad()
    def ema_update(self):
        """ddp_model.py:26-30 / :93-97 / :178-182 as two multi-tensor launches instead of a Python loop over ~600
        parameters."""
        decay = min(1 - 1 / (self.step + 1), self.alpha)
        ema = [p.data for p in self.ema_model.parameters()]
        cur = [p.data for p in self.model.parameters()]
        if ema:
            torch._foreach_mul_(ema, decay)
            torch._foreach_add_(ema, cur, alpha=1 - decay)
        self.step += 1

    def _student(self, train_l_image, train_u_aug_image):
        pred_l, rep_l = self.model(train_l_image)
        pred_l_large = F.interpolate(pred_l, size=train_l_image.shape[2:], mode='bilinear', align_corners=True)
        pred_u, rep_u = self.model(train_u_aug_image)
        pred_u_large = F.interpolate(pred_u, size=train_l_image.shape[2:], mode='bilinear', align_corners=True)
        return pred_l_large, pred_u_large, torch.cat((rep_l, rep_u)), torch.cat((pred_l, pred_u))

    def _aug_cfg(self):
        d = self.config['Dataset']
        return d['crop_size'], d['scale_size'], d['mix_mode']


class Model_ori_pseudo(_TeacherStudent):
    """ddp_model.py:8-70."""

    def __init__(self, base_encoder, num_classes=21, output_dim=256, ema_alpha=0.99, config=None) -> None:
        super().__init__(base_encoder, num_classes, output_dim, ema_alpha, config)

    def forward(self, train_l_image, train_u_image):
        crop, scale, mix = self._aug_cfg()
        with torch.no_grad():
            pred_u, _ = self.ema_model(train_u_image)
            # the raw up-sampled logits are part of the return tuple (:70), so they are materialised here as well
            pred_u_large_raw = F.interpolate(pred_u, size=train_u_image.shape[2:], mode='bilinear', align_corners=True)
            pseudo_logits, pseudo_labels = ops.cls_pseudo_label(pred_u, train_u_image.shape[2:])            # :36-37
            u_img, u_label, u_logits = _hook("batch_transform")(train_u_image, pseudo_labels, pseudo_logits, crop_size=crop,
                                                                scale_size=scale, augmentation=False)
            u_img, u_label, u_logits = _hook("generate_cut_gather")(u_img, u_label, u_logits, mode=mix)
            u_img, u_label, u_logits = _hook("batch_transform")(u_img, u_label, u_logits, crop_size=crop,
                                                                scale_size=(1.0, 1.0), augmentation=True)
        pred_l_large, pred_u_large, rep_all, pred_all = self._student(train_l_image, u_img)
        return pred_l_large, pred_u_large, u_label, u_logits, rep_all, pred_all, pred_u_large_raw


class Model_mix(_TeacherStudent):
    """ddp_model.py:73-156."""

    def __init__(self, base_encoder, num_classes=21, output_dim=256, ema_alpha=0.99, config=None, temp=0.25) -> None:
        super().__init__(base_encoder, num_classes, output_dim, ema_alpha, config)
        self.temp = temp

    def forward(self, train_l_image, train_u_image, prototypes):
        crop, scale, mix = self._aug_cfg()
        with torch.no_grad():
            self.ema_model(train_l_image)                       # :102 (kept: it updates the teacher's BN statistics)
            pred_u, rep_u = self.ema_model(train_u_image)
            o = ops.pseudo_labels(rep_u, pred_u, prototypes, self.temp, train_u_image.shape[2:], fuse="mix")   # :104-118
            u_img, u_label, u_lc, u_lr = _hook("batch_transform_2")(train_u_image, o["fused"], o["conf_cls"], o["conf_rep"],
                                                                    crop_size=crop, scale_size=scale, augmentation=False)
            u_img, u_label, u_lc, u_lr = _hook("generate_cut_gather_2")(u_img, u_label, u_lc, u_lr, mode=mix)
            u_img, u_label, u_lc, u_lr = _hook("batch_transform_2")(u_img, u_label, u_lc, u_lr, crop_size=crop,
                                                                    scale_size=(1.0, 1.0), augmentation=True)
        pred_l_large, pred_u_large, rep_all, _ = self._student(train_l_image, u_img)
        prob_all = ops.proto_softmax_sim(rep_all, prototypes, self.temp)                                     # :147-154
        return pred_l_large, pred_u_large, u_label, u_lc, u_lr, rep_all, prob_all


class Model_cross(_TeacherStudent):
    """ddp_model.py:158-239."""

    def __init__(self, base_encoder, num_classes=21, output_dim=256, ema_alpha=0.99, config=None, temp=0.1) -> None:
        super().__init__(base_encoder, num_classes, output_dim, ema_alpha, config)
        self.temp = temp

    def forward(self, train_l_image, train_u_image, prototypes):
        crop, scale, mix = self._aug_cfg()
        with torch.no_grad():
            self.ema_model(train_l_image)                       # :187
            pred_u, rep_u = self.ema_model(train_u_image)
            o = ops.pseudo_labels(rep_u, pred_u, prototypes, self.temp, train_u_image.shape[2:], fuse="none")  # :189-199
            r = _hook("batch_transform_3")(train_u_image, o["label_cls"], o["label_rep"], o["conf_cls"], o["conf_rep"],
                                           crop_size=crop, scale_size=scale, augmentation=False)
            r = _hook("generate_cut_gather_3")(*r, mode=mix)
            u_img, u_label_cls, u_label_rep, u_lc, u_lr = _hook("batch_transform_3")(*r, crop_size=crop, scale_size=(1.0, 1.0),
                                                                                     augmentation=True)
        pred_l_large, pred_u_large, rep_all, _ = self._student(train_l_image, u_img)
        prob_all = ops.proto_softmax_sim(rep_all, prototypes, self.temp)                                     # :230-237
        return pred_l_large, pred_u_large, u_label_cls, u_label_rep, u_lc, u_lr, rep_all, prob_all
