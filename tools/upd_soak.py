"""Soak of the PPO update kernels over random minibatch shapes: gradients against torch fp32 autograd (cosine, max error),
run-to-run and kernel-variant (persistent / one-tile-per-CTA) bitwise equality, sticky error word.
python tools/upd_soak.py [iterations] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from metabo_b200 import ops
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 40
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = torch.device("cuda:0")
rng = np.random.RandomState(seed)
torch.backends.cuda.matmul.allow_tf32 = False
worst_cos, worst_err = 1.0, 0.0
t0 = time.time()
for it in range(iters):
    E = int(rng.choice([1, 2, 5, 17, 60, 119]))
    M = int(rng.choice([2, 31, 128, 129, 966, 2000, 5000]))
    if E * M > 300000:
        M = 966
    H = int(rng.choice([8, 20, 64, 100, 200, 207]))
    Fs = int(rng.choice([4, 6, 7, 8]))
    cols = sorted(rng.choice(Fs, size=int(rng.randint(1, Fs + 1)), replace=False).tolist())
    F = len(cols)
    g = torch.Generator().manual_seed(int(rng.randint(1 << 30)))
    dims = [F, H, H, H, H, 1]
    sc = 1.5 / np.sqrt(H)
    W = [(torch.randn(dims[i + 1], dims[i], generator=g) * (0.5 if i == 0 else sc)).to(dev).requires_grad_() for i in range(5)]
    b = [(torch.randn(dims[i + 1], generator=g) * 0.05).to(dev).requires_grad_() for i in range(5)]
    st = torch.randn(E, M, Fs, generator=g).to(dev)
    ac = torch.randint(0, M, (E,), generator=g).to(dev)
    adv = torch.randn(E, generator=g).to(dev)
    h = st[:, :, cols].reshape(E * M, F)
    for l in range(4):
        h = torch.relu(h @ W[l].t() + b[l])
    logits = (h @ W[4].t() + b[4]).reshape(E, M)
    norm = logits - torch.logsumexp(logits, -1, keepdim=True)
    lp = norm.gather(-1, ac[:, None]).squeeze(-1)
    lpo = (lp.detach() + 0.2 * torch.randn(E, generator=g).to(dev)).contiguous()
    ent = -(norm * norm.exp()).sum(-1)
    ratio = torch.exp(lp - lpo)
    loss = -torch.min(ratio * adv, ratio.clamp(0.85, 1.15) * adv).sum() / E - 0.01 * ent.sum() / E
    loss.backward()
    ref = torch.cat([p.grad.reshape(-1) for p in W + b[:4]]).double()
    pg = ops.PPOPolicyGrad(E, M, Fs, cols, H, dev)
    Wd, bd = [w.detach().contiguous() for w in W], [x.detach().contiguous() for x in b]
    outs = []
    for dgr, ts in (("1", "1"), ("0", "0"), ("1", "1")):
        os.environ["MBO_UPD_PERSIST_DGRAD"], os.environ["MBO_UPD_FWD_TS"] = dgr, ts
        dW, db = [torch.full_like(w, 9.0) for w in Wd], [torch.full_like(x, 9.0) for x in bd]
        terms = pg(st, Wd, bd, ac.to(torch.int32), adv, lpo, 0.15, 0.01, E, dW, db).clone()
        torch.cuda.synchronize(); pg.check()
        outs.append(torch.cat([t.reshape(-1) for t in dW + db + [terms]]))
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2]), ("variant / run-to-run difference", it, E, M, H, cols)
    got = outs[0][:ref.numel()].double()
    assert torch.isfinite(outs[0]).all(), ("non-finite", it, E, M, H, cols)
    if float(ref.abs().max()) > 0:
        cos = float(torch.nn.functional.cosine_similarity(got, ref, dim=0))
        err = float((got - ref).abs().max() / ref.abs().max())
        worst_cos, worst_err = min(worst_cos, cos), max(worst_err, err)
        assert cos > 0.999, ("gradient direction", it, E, M, H, cols, cos, err)
    lp_err = float((outs[0][-4 * E:].view(E, 4)[:, 0] - lp.detach()).abs().max())
    assert lp_err <= 2e-4 * max(1.0, float(logits.detach().abs().max())), ("log-prob", it, E, M, H, cols, lp_err)
    del pg
print("update soak ok: %d shapes, %.1f s, worst cosine %.6f, worst max-norm gradient error %.2e" % (iters, time.time() - t0, worst_cos, worst_err))
