"""Small driver for ncu captures of the tcgen05 NeuralAF kernel (one mode, B x M candidates)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metabo_b200 import ops
mode = {"tf32": 1, "bf16x3": 2, "fp16": 3}[sys.argv[1] if len(sys.argv) > 1 else "tf32"]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
M = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
dev = torch.device("cuda:0")
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "weights_hm3_1988.npz"))
W = [torch.as_tensor(z["policy_net.fc.%d.weight" % i], device=dev) for i in range(5)]
b = [torch.as_tensor(z["policy_net.fc.%d.bias" % i], device=dev) for i in range(5)]
pol = ops.PolicyHandle(); pol.set_weights(W, b, "relu")
g = torch.Generator(device=dev); g.manual_seed(0)
cs = ops.CandidateSet(3, tail=torch.rand((M, 3), generator=g, dtype=torch.float64, device=dev))
mean = torch.randn((B, M), generator=g, device=dev); std = torch.rand((B, M), generator=g, device=dev)
epi = torch.zeros(B, 4, device=dev); epi[:, 2] = 15; epi[:, 3] = 30
topk = len(sys.argv) > 4 and sys.argv[4] == "topk"      # the AF round's search phases: selection in the epilogue, no logits written
lg = None
for _ in range(3):
    if topk:
        idx, val = ops.af_topk(pol, [0, 1, 2, 3, 4, 12, 13], cs, mean, std, epi, B, 5, mode=mode)
    else:
        lg = ops.af_logits(pol, [0, 1, 2, 3, 4, 12, 13], cs, mean, std, epi, B, mode=mode, logits=lg)
torch.cuda.synchronize(); pol.check()
print("ok", float(val.abs().max()) if topk else float(lg.abs().max()))
