import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metabo_b200 import ops
dev = torch.device("cuda:0")
B, M, D, T, n = 1024, 10000, 3, 30, 30
g = torch.Generator(device=dev); g.manual_seed(0)
cs = ops.CandidateSet(D, tail=torch.rand((M, D), generator=g, dtype=torch.float64, device=dev))
X = torch.rand((B, T, D), generator=g, dtype=torch.float64, device=dev)
Y = torch.randn((B, T), generator=g, dtype=torch.float64, device=dev)
nn = torch.full((B,), n, dtype=torch.int32, device=dev)
ls = torch.tensor([[0.716, 0.298, 0.186]] * B, dtype=torch.float64, device=dev)
st, _ = ops.gp_fit(nn, X, Y, ls, torch.full((B,), 0.83, dtype=torch.float64, device=dev), torch.full((B,), 1e-6, dtype=torch.float64, device=dev))
mean = torch.empty((B, M), device=dev); std = torch.empty((B, M), device=dev)
for _ in range(3):
    ops.gp_posterior(st, T, cs, B, out_dtype=torch.float32, n_active_max=n, mean=mean, std=std)
torch.cuda.synchronize(); print("ok")
