"""ncu target: a few launches of the rep pass (tc and fma paths) on the V321 student / teacher maps."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import css_b200  # noqa: E402
from css_b200 import _lib, synth  # noqa: E402

lib = _lib.load()
d = synth.student_batch(16, 21, 81, 81, seed=1)
rep = d["rep"].cuda()
rep_u = rep[:8].contiguous()
protos = synth.warm_prototypes(21, seed=2).cuda()
for flag in (1, 0):
    lib.css_set_rep_pass_path(flag)
    for _ in range(3):
        css_b200.ops.cos_sim_map(rep_u, protos)
        css_b200.ops.proto_softmax_sim(rep, protos, 0.5)
torch.cuda.synchronize()
