"""Timing probe: fused GP+MLP kernel vs the two-kernel path for several n (B=1024, M=10000, HM3 weights)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metabo_b200 import ops, _lib as L
dev = torch.device("cuda:0")
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
z = np.load(os.path.join(root, "tests", "golden", "weights_hm3_1988.npz"))
W = [torch.as_tensor(z["policy_net.fc.%d.weight" % i], device=dev) for i in range(5)]
b = [torch.as_tensor(z["policy_net.fc.%d.bias" % i], device=dev) for i in range(5)]
pol = ops.PolicyHandle(); pol.set_weights(W, b, "relu")
B, M, D, T = 1024, 10000, 3, 30
g = torch.Generator(device=dev); g.manual_seed(0)
cs = ops.CandidateSet(D, tail=torch.rand((M, D), generator=g, dtype=torch.float64, device=dev))
epi = torch.zeros(B, 4, device=dev); epi[:, 2] = 15; epi[:, 3] = 30
codes = [0, 1, 2, 3, 4, 12, 13]
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps
for n in (0, 8, 16, 30):
    X = torch.rand((B, T, D), generator=g, dtype=torch.float64, device=dev)
    Y = torch.randn((B, T), generator=g, dtype=torch.float64, device=dev)
    nn = torch.full((B,), n, dtype=torch.int32, device=dev)
    ls = torch.tensor([[0.716, 0.298, 0.186]] * B, dtype=torch.float64, device=dev)
    st, _ = ops.gp_fit(nn, X, Y, ls, torch.full((B,), 0.83, dtype=torch.float64, device=dev),
                       torch.full((B,), 1e-6, dtype=torch.float64, device=dev))
    lg = torch.empty((B, M), device=dev)
    mean = torch.empty((B, M), device=dev); std = torch.empty((B, M), device=dev)
    for name, mode in (("tf32", 1), ("bf16x3", 2)):
        tf = timeit(lambda: ops.af_eval_fused(pol, codes, cs, st, T, max(n, 1), epi, B, mode=mode, logits=lg))
        tg = timeit(lambda: ops.gp_posterior(st, T, cs, B, out_dtype=torch.float32, n_active_max=max(n, 1), mean=mean, std=std))
        tm = timeit(lambda: ops.af_logits(pol, codes, cs, mean, std, epi, B, mode=mode, logits=lg))
        pol.check()
        print("n=%2d %-6s fused %.3f ms | gp %.3f + mlp %.3f = %.3f ms" % (n, name, tf, tg, tm, tg + tm), flush=True)
