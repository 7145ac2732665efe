import sys, os, numpy as np, torch
sys.path.insert(0, "/root/repo")
from metabo_b200 import ops
from oracle import metabo_oracle as O
dev = torch.device("cuda:0")
pw = O.PolicyWeights.from_npz("tests/golden/weights_hm3_1988.npz")
layers = pw.layers("value_net")
vh = ops.PolicyHandle(); vh.set_weights([w.cuda() for w, _ in layers], [b.cuda() for _, b in layers], "relu")
for B in (60, 1024, 4096):
    tT = torch.rand(B, 2, device=dev) * 30
    v = ops.value_net(vh, tT); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20): ops.value_net(vh, tT)
    e.record(); torch.cuda.synchronize()
    h = tT.double().cpu()
    for i, (w, b) in enumerate(layers):
        h = h @ w.double().T + b.double()
        if i < len(layers) - 1: h = torch.relu(h)
    print("B=%d  %.1f us  max err %.2e" % (B, s.elapsed_time(e) * 50, float((v.cpu().double() - h[:, 0]).abs().max())))
