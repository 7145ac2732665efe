"""GP posterior timing by n (B=1024, M=16384, D=3), narrow vs wide kernel (MBO_GP_WIDE_FROM)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metabo_b200 import ops
dev = torch.device("cuda:0")
B, M, D = 1024, 16384, 3
g = torch.Generator(device=dev); g.manual_seed(0)
cs = ops.CandidateSet(D, tail=torch.rand((M, D), generator=g, dtype=torch.float64, device=dev))
mean = torch.empty((B, M), device=dev); std = torch.empty((B, M), device=dev)
for n in (16, 30, 32, 48, 64, 100, 128):
    X = torch.rand((B, n, D), generator=g, dtype=torch.float64, device=dev)
    Y = torch.randn((B, n), generator=g, dtype=torch.float64, device=dev)
    st, _ = ops.gp_fit(torch.full((B,), n, dtype=torch.int32, device=dev), X, Y, torch.full((B, D), 0.3, dtype=torch.float64, device=dev),
                       torch.ones(B, dtype=torch.float64, device=dev), torch.full((B,), 1e-4, dtype=torch.float64, device=dev))
    f = lambda: ops.gp_posterior(st, n, cs, B, out_dtype=torch.float32, n_active_max=n, mean=mean, std=std)
    f(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); f(); f(); e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 2
    print("n=%3d  %.3f ms  %.2f TFLOP/s (alg)  checksum %.6f" % (n, ms, B * M * (n * n + n * 17) / ms / 1e9, float(mean.double().sum() + std.double().sum())), flush=True)
