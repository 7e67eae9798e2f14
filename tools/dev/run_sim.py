import ctypes, os, subprocess, sys, torch
HERE = os.path.dirname(os.path.abspath(__file__))
so = os.path.join(HERE, "libdevsim.so")
lib = ctypes.CDLL(so)
P = ctypes.c_void_p
lib.dev_sim.argtypes = [ctypes.c_int, P, P, ctypes.c_int, ctypes.c_int, ctypes.c_int, P, P]
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import css_b200
dev = torch.device("cuda")
C, h, w = 21, 81, 81
protos = torch.randn(C, 256, device=dev)
for B in (8, 16):
    pool = [torch.randn(B, 256, h, w, device=dev) for _ in range(4)]
    ref = css_b200.ops.cos_sim_map(pool[0], protos)
    # scratch as the product prepares it
    pn = torch.nn.functional.normalize(protos, dim=-1)
    scratch = torch.zeros(256, 32, device=dev); scratch[:, :C] = pn.t()
    out = torch.empty(B, C, h, w, device=dev)
    N = B * h * w
    for v in range(12):
        def fn(i):
            rc = lib.dev_sim(v, P(pool[i % 4].data_ptr()), P(scratch.data_ptr()), h * w, N, C, P(out.data_ptr()), P(torch.cuda.current_stream().cuda_stream))
            assert rc == 0, rc
        fn(0); torch.cuda.synchronize()
        err = (out - ref).abs().max().item()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(20): fn(i)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        print(f"B={B} variant {v:2d}: {us:7.1f} us  {B*256*h*w*4/us/1e3:7.1f} GB/s  maxerr {err:.2e}")
