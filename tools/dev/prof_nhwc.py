"""ncu target: the channels-last rep pass (TMA + tcgen05) on the V321 student / teacher maps."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import css_b200  # noqa: E402
from css_b200 import synth  # noqa: E402

d = synth.student_batch(16, 21, 81, 81, seed=1)
rep = d["rep"].cuda().contiguous(memory_format=torch.channels_last)
rep_u = rep[:8].contiguous(memory_format=torch.channels_last)
protos = synth.warm_prototypes(21, seed=2).cuda()
for _ in range(3):
    css_b200.ops.cos_sim_map(rep_u, protos)
    css_b200.ops.proto_softmax_sim(rep, protos, 0.5)
torch.cuda.synchronize()
