import os, sys, torch
sys.path.insert(0, os.getcwd())
from metabo_b200 import ops
dev = torch.device("cuda:0")
E, M, F, H = 60, 966, 6, 200
g = torch.Generator().manual_seed(0)
dims = [F, H, H, H, H, 1]
W = [(torch.randn(dims[i + 1], dims[i], generator=g) * 0.08).to(dev) for i in range(5)]
b = [(torch.randn(dims[i + 1], generator=g) * 0.05).to(dev) for i in range(5)]
st = torch.randn(E, M, F, generator=g).to(dev)
ac = torch.randint(0, M, (E,), generator=g).to(dev).to(torch.int32)
adv = torch.randn(E, generator=g).to(dev)
pg = ops.PPOPolicyGrad(E, M, F, list(range(F)), H, dev)
dW, db = [torch.zeros_like(w) for w in W], [torch.zeros_like(x) for x in b]
lpo, _ = pg.logp_entropy(st, W, b, ac)
for _ in range(3):
    pg(st, W, b, ac, adv, lpo, 0.15, 0.01, E, dW, db)
torch.cuda.synchronize(); pg.check(); print("ok")
