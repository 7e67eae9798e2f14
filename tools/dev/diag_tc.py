"""Diagnostic for the tcgen05 rep pass: which B channel is paired with which A channel, and which pixel lands in which TMEM lane."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from css_b200 import _lib, ops  # noqa: E402


def raw(rep, protos):
    lib = _lib.load()
    B, D, h, w = rep.shape
    C = protos.shape[0]
    out = torch.empty(B, C, h, w, device="cuda")
    scratch = ops._proto_scratch(rep.device)
    _lib.check(lib.css_rep_pass(_lib.ptr(rep), 0, _lib.ptr(protos), _lib.ptr(scratch), B, C, D, h, w, 2, 1.0, _lib.ptr(out), None, None,
                                _lib.stream_ptr()), "css_rep_pass")
    torch.cuda.synchronize()
    return out


def main():
    lib = _lib.load()
    lib.css_set_rep_pass_path(1)
    h, w, C = 8, 16, 4                                     # one 128-pixel tile
    protos = (torch.arange(256, dtype=torch.float32) + 1).repeat(C, 1).cuda()
    norm = float(protos[0].norm())
    pair = []
    for k0 in range(256):
        rep = torch.zeros(1, 256, h, w, device="cuda")
        rep[:, k0] = 1
        o = raw(rep, protos)[0, 0].reshape(-1) * norm - 1       # per pixel: paired B channel
        vals = o.cpu().numpy()
        pair.append((float(vals.min()), float(vals.max())))
    pair = np.array(pair)
    print("A channel -> paired B channel (min/max over pixels), first 40 and any non-identity:")
    for k0 in range(256):
        if k0 < 40 or abs(pair[k0, 0] - k0) > 0.01 or abs(pair[k0, 1] - k0) > 0.01:
            print(f"  k={k0}: {pair[k0, 0]:.3f} .. {pair[k0, 1]:.3f}")
    # pixel -> lane: channel 0 carries pixel id + 1, prototypes all ones -> raw = (pixel' + 1) / 16
    protos1 = torch.ones(C, 256).cuda()
    rep = torch.zeros(1, 256, h, w, device="cuda")
    rep[0, 0] = (torch.arange(h * w, dtype=torch.float32) + 1).reshape(h, w).cuda()
    o = (raw(rep, protos1)[0, 0].reshape(-1) * 16 - 1).cpu().numpy()
    print("output pixel p shows input pixel:", np.round(o, 2).tolist())
    # same with channel 9
    rep = torch.zeros(1, 256, h, w, device="cuda")
    rep[0, 9] = (torch.arange(h * w, dtype=torch.float32) + 1).reshape(h, w).cuda()
    o = (raw(rep, protos1)[0, 0].reshape(-1) * 16 - 1).cpu().numpy()
    print("channel 9: output pixel p shows input pixel:", np.round(o, 2).tolist())
    # class mapping: prototype n = (n + 1) * ones, rep = ones on channel 0
    protos2 = torch.stack([torch.full((256,), float(n + 1)) for n in range(C)]).cuda()     # normalised: all 1/16 -> use distinct channels instead
    protos2 = torch.zeros(C, 256)
    for n in range(C):
        protos2[n, n] = 1.0
    rep = torch.zeros(1, 256, h, w, device="cuda")
    for n in range(C):
        rep[0, n] = float(n + 1)
    o = raw(rep, protos2.cuda())[0].reshape(C, -1)[:, :4].cpu().numpy()
    print("class n column (expect n+1):", o.tolist())
    lib.css_set_rep_pass_path(-1)


if __name__ == "__main__":
    main()
