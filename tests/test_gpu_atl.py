"""GPU parity of the fused Attention_Threshold_Loss against the golden bundles recorded from the live
reference (loss.py:48-64, forward value + autograd gradient) and against the numpy oracle at crop size.  rel 1e-4."""
import numpy as np
import pytest
import torch

from oracle import css_oracle as O
from tests.helpers import load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["atl_c21", "atl_c19"])
def test_atl_against_reference_golden(name):
    import css_b200
    g = load_golden(name)
    crit = css_b200.Attention_Threshold_Loss(strong_threshold=float(g["thr"])).cuda()
    pred = torch.from_numpy(g["pred"]).cuda().requires_grad_(True)
    loss = crit(pred, torch.from_numpy(g["label"].astype(np.int64)).cuda(), torch.from_numpy(g["conf"]).cuda())
    (loss * float(g["grad_scale"])).backward()
    assert loss.dim() == 0
    np.testing.assert_allclose(loss.item(), g["loss"], rtol=1e-4)
    gg = pred.grad.cpu().numpy()
    assert np.array_equal(gg != 0, g["grad"] != 0)
    np.testing.assert_allclose(gg, g["grad"], rtol=1e-4, atol=1e-9)


def test_atl_against_oracle_crop_size_and_edge_cases():
    import css_b200
    from css_b200 import synth
    B, C, H, W = 4, 21, 161, 161
    gen = synth._gen(5)
    cls = synth.class_map(B, C, H, W, gen, ignore_frac=0.2)
    cls[1] = -1                                               # a fully ignored image: its weight is 0/0 = nan, never selected
    pred = synth.logits_for(cls, C, gen)
    conf = torch.rand(B, H, W, generator=gen)
    crit = css_b200.Attention_Threshold_Loss(strong_threshold=0.97).cuda()
    p = pred.cuda().requires_grad_(True)
    loss = crit(p, cls.cuda(), conf.cuda())
    loss.backward()
    l_or, g_or = O.attention_threshold_loss(pred.numpy(), cls.numpy(), conf.numpy(), 0.97)
    assert np.isfinite(loss.item())
    np.testing.assert_allclose(loss.item(), l_or, rtol=1e-4)
    gg = p.grad.cpu().numpy()
    assert not gg[1].any() and np.isfinite(gg).all()
    np.testing.assert_allclose(gg, g_or, rtol=1e-4, atol=1e-10)
    # everything ignored -> mean of an empty selection is nan (as torch.mean of an empty tensor)
    l2 = crit(pred.cuda(), torch.full_like(cls, -1).cuda(), conf.cuda())
    assert np.isnan(l2.item())
    with pytest.raises(RuntimeError, match="CUDA"):
        crit(pred, cls, conf)
