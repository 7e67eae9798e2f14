import ctypes, os, sys, torch
HERE = os.path.dirname(os.path.abspath(__file__))
lib = ctypes.CDLL(os.path.join(HERE, "libdevsim3.so"))
P = ctypes.c_void_p
lib.dev_sim3.argtypes = [ctypes.c_int, P, P, ctypes.c_int, ctypes.c_int, ctypes.c_int, P, P]
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import css_b200
dev = torch.device("cuda")
C = 21
protos = torch.randn(C, 256, device=dev)
for (B, h, w) in ((8, 81, 81), (16, 81, 81), (4, 193, 193)):
    pool = [torch.randn(B, 256, h, w, device=dev) for _ in range(4)]
    ref = css_b200.ops.cos_sim_map(pool[0], protos)
    pn = torch.nn.functional.normalize(protos, dim=-1)
    scratch = torch.zeros(256, 32, device=dev); scratch[:, :C] = pn.t()
    out = torch.empty(B, C, h, w, device=dev)
    N = B * h * w
    def timeit(fn):
        fn(0); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(20): fn(i)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 20 * 1e3
    us = timeit(lambda i: css_b200.ops.cos_sim_map(pool[i % 4], protos))
    print(f"B={B} {h}x{w} product cos_sim_map (incl. proto_prep): {us:7.1f} us")
    for v in (0, 14, 15, 16, 17, 18, 19):
        def fn(i):
            rc = lib.dev_sim3(v, P(pool[i % 4].data_ptr()), P(scratch.data_ptr()), h * w, N, C, P(out.data_ptr()), P(torch.cuda.current_stream().cuda_stream))
            assert rc == 0, rc
        out.zero_()
        fn(0); torch.cuda.synchronize()
        err = (out - ref).abs().max().item()
        us = timeit(fn)
        print(f"B={B} {h}x{w} variant {v:2d}: {us:7.1f} us  {B*256*h*w*4/us/1e3:7.1f} GB/s  maxerr {err:.2e}")
