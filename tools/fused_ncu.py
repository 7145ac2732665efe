import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metabo_b200 import ops
dev = torch.device("cuda:0")
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
z = np.load(os.path.join(root, "tests", "golden", "weights_hm3_1988.npz"))
W = [torch.as_tensor(z["policy_net.fc.%d.weight" % i], device=dev) for i in range(5)]
b = [torch.as_tensor(z["policy_net.fc.%d.bias" % i], device=dev) for i in range(5)]
pol = ops.PolicyHandle(); pol.set_weights(W, b, "relu")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
B, M, D, T = 256, 10000, 3, 30
g = torch.Generator(device=dev); g.manual_seed(0)
cs = ops.CandidateSet(D, tail=torch.rand((M, D), generator=g, dtype=torch.float64, device=dev))
epi = torch.zeros(B, 4, device=dev); epi[:, 2] = 15; epi[:, 3] = 30
X = torch.rand((B, T, D), generator=g, dtype=torch.float64, device=dev)
Y = torch.randn((B, T), generator=g, dtype=torch.float64, device=dev)
nn = torch.full((B,), n, dtype=torch.int32, device=dev)
ls = torch.tensor([[0.716, 0.298, 0.186]] * B, dtype=torch.float64, device=dev)
st, _ = ops.gp_fit(nn, X, Y, ls, torch.full((B,), 0.83, dtype=torch.float64, device=dev), torch.full((B,), 1e-6, dtype=torch.float64, device=dev))
lg = torch.empty((B, M), device=dev)
for _ in range(3):
    ops.af_eval_fused(pol, [0, 1, 2, 3, 4, 12, 13], cs, st, T, max(n, 1), epi, B, mode=1, logits=lg)
torch.cuda.synchronize(); pol.check(); print("ok")
