"""Probe of the hand-written PPO policy-gradient kernels (mbo_ppo_policy_grad): per-stage comparison with torch
fp32 autograd (activations, logits, dlogits, dZ chain, every gradient) and timing against the autograd step.
python tools/ppo_update_probe.py [E] [M] [H] [F]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from metabo_b200 import ops

E = int(sys.argv[1]) if len(sys.argv) > 1 else 60
M = int(sys.argv[2]) if len(sys.argv) > 2 else 966
H = int(sys.argv[3]) if len(sys.argv) > 3 else 200
F = int(sys.argv[4]) if len(sys.argv) > 4 else 6
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
g = torch.Generator(device="cpu").manual_seed(0)
dims = [F, H, H, H, H, 1]
scale = float(os.environ.get("PROBE_WSCALE", 0.08))
Ws = [(torch.randn(dims[i + 1], dims[i], generator=g) * scale).to(dev).requires_grad_() for i in range(5)]
bs = [(torch.randn(dims[i + 1], generator=g) * 0.05).to(dev).requires_grad_() for i in range(5)]
states = torch.randn(E, M, F, generator=g).to(dev)
actions = torch.randint(0, M, (E,), generator=g).to(dev)
advs = torch.randn(E, generator=g).to(dev)
eps, entc = 0.15, 0.01


def torch_step(keep=False):
    hs = []
    h = states.reshape(E * M, F)
    for l in range(4):
        h = torch.relu(h @ Ws[l].t() + bs[l])
        if keep:
            h.retain_grad()
        hs.append(h)
    logits = (h @ Ws[4].t() + bs[4]).reshape(E, M)
    if keep:
        logits.retain_grad()
    norm = logits - torch.logsumexp(logits, -1, keepdim=True)
    lp = norm.gather(-1, actions[:, None]).squeeze(-1)
    ent = -(norm * norm.exp()).sum(-1)
    return hs, logits, lp, ent


with torch.no_grad():
    _, _, lp0, _ = torch_step()
logp_old = (lp0 + 0.2 * torch.randn(E, generator=g).to(dev)).contiguous()    # some ratios outside the clip range


def torch_loss():
    hs, logits, lp, ent = torch_step(keep=True)
    ratio = torch.exp(lp - logp_old)
    clipped = ratio.clamp(1 - eps, 1 + eps)
    loss_ppo = -torch.min(ratio * advs, clipped * advs).sum() / E
    loss_ent = -ent.sum() / E
    return hs, logits, lp, ent, ratio, loss_ppo + entc * loss_ent


hs, logits, lp, ent, ratio, loss = torch_loss()
loss.backward()
torch.cuda.synchronize()

pg = ops.PPOPolicyGrad(E, M, F, list(range(F)), H, dev)
dW = [torch.zeros_like(w) for w in Ws]
db = [torch.zeros_like(b) for b in bs]
Wd = [w.detach().contiguous() for w in Ws]
bd = [b.detach().contiguous() for b in bs]
a32 = actions.to(torch.int32)
terms = pg(states, Wd, bd, a32, advs, logp_old, eps, entc, E, dW, db)
torch.cuda.synchronize()
pg.check()
v = pg.debug_views()


def rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


print("shape E=%d M=%d H=%d F=%d rows=%d" % (E, M, H, F, E * M))
for l in range(4):
    got = v["h%d" % (l + 1)]
    print("h%d   rel err %.3e   ones col ok %s   pad cols zero %s" % (
        l + 1, rel(got[:, :H], hs[l].detach()), bool((got[:, H] == 1).all()), bool((got[:, H + 1:] == 0).all())))
print("logits rel err %.3e" % rel(v["logits"], logits.detach().reshape(-1)))
print("logp   rel err %.3e   entropy %.3e   ratio %.3e" % (rel(terms[:, 0], lp.detach()), rel(terms[:, 1], ent.detach()),
                                                          rel(terms[:, 3], ratio.detach())))
print("dlogits rel err %.3e" % rel(v["dlogits"], logits.grad.reshape(-1)))
for l in range(3, -1, -1):
    # dZ_l = dh_l * [h_l > 0]
    ref = hs[l].grad * (hs[l].detach() > 0)
    got = v["dz%d" % (l + 1)] * (v["h1"] > 0) if l == 0 else v["dz%d" % (l + 1)]     # l = 0: the kernel stores dH1
    print("dz%d  rel err %.3e   pad cols zero %s" % (l + 1, rel(got[:, :H], ref), bool((got[:, H:] == 0).all())))
for l in range(5):
    print("dW%d rel err %.3e   db%d rel err %.3e   (max |dW| %.3e)" % (l, rel(dW[l], Ws[l].grad), l, rel(db[l], bs[l].grad),
                                                                  float(Ws[l].grad.abs().max())))
# The tf32 forward flips the ReLU state of pre-activations that are within ~1e-4 of zero; such an element changes dZ by
# its full value.  Separate that from kernel errors: redo the backward in fp64 WITH THE KERNEL'S OWN masks.
with torch.no_grad():
    flips = [int(((v["h%d" % (l + 1)][:, :H] > 0) != (hs[l] > 0)).sum()) for l in range(4)]
    print("ReLU flips vs fp32 forward per layer:", flips, "of", E * M * H)
    dl64 = v["dlogits"].double()
    myh = [v["h%d" % (l + 1)][:, :H].double() for l in range(4)]
    W64 = [w.detach().double() for w in Ws]
    dz = (dl64[:, None] * W64[4]) * (myh[3] > 0)
    x64 = states.reshape(E * M, F).double()
    for l in range(3, -1, -1):
        gdz = v["dz%d" % (l + 1)] * (v["h1"] > 0) if l == 0 else v["dz%d" % (l + 1)]
        print("  own-mask reference: dz%d rel err %.3e" % (l + 1, rel(gdz[:, :H].double(), dz)), end="")
        inp = myh[l - 1] if l > 0 else x64
        print("   dW%d %.3e   db%d %.3e" % (l, rel(dW[l].double(), dz.t() @ inp), l, rel(db[l].double(), dz.sum(0))))
        if l > 0:
            dz = (dz @ W64[l]) * (myh[l - 1] > 0)
    # what torch's own TF32 autograd does on the same problem
    torch.backends.cuda.matmul.allow_tf32 = True
ref_grads = [w.grad.clone() for w in Ws]
for p_ in Ws + bs:
    p_.grad = None
_, _, _, _, _, loss_t = torch_loss()
loss_t.backward()
print("torch TF32 autograd vs torch fp32 autograd: " + "  ".join("dW%d %.3e" % (l, rel(Ws[l].grad, ref_grads[l])) for l in range(5)))
torch.backends.cuda.matmul.allow_tf32 = False
for l in range(5):
    Ws[l].grad = ref_grads[l]
cos = torch.nn.functional.cosine_similarity(torch.cat([g.reshape(-1) for g in dW]), torch.cat([g.reshape(-1) for g in ref_grads]), dim=0)
print("cosine(native grad, fp32 autograd grad) = %.6f" % float(cos))
if os.environ.get("PROBE_WG"):
    # split-K partial of CTA (split 0, layer 1): rows [0, per*32) of dZ2^T [h1, 1]
    nk = (E * M + 31) // 32
    per = (nk + 48) // 49
    r1 = min(E * M, per * 32)
    for variant in (0,):
        os.environ["MBO_WG_VARIANT"] = str(variant)
        pg(states, Wd, bd, a32, advs, logp_old, eps, entc, E, dW, db)
        torch.cuda.synchronize()
        try:
            pg.check()
        except Exception as ex:
            print("variant", variant, "error:", ex)
        vv = pg.debug_views()
        exp = vv["dz2"][:r1].double().t() @ vv["h1"][:r1].double()      # [208(o), 208(i)]
        got = vv["part_big"][0, 0].double()
        print("variant %d: partial max |got| %.3e  max |exp| %.3e  rel err (o<H) %.3e" % (
            variant, float(got.abs().max()), float(exp.abs().max()), rel(got[:H], exp[:H])))
        print(" got[:2,:6]\n", got[:2, :6].cpu().numpy(), "\n exp[:2,:6]\n", exp[:2, :6].cpu().numpy())
        print(" nonzero fraction of all partials %.4f" % float((vv["part_big"] != 0).float().mean()))
        # hypothesis tests: transposed result, K subsets
        print(" rel err vs exp^T %.3e" % rel(got[:H, :H], exp[:H, :H].t()))
        print(" dW1 rel err %.3e" % rel(dW[1], Ws[1].grad))
    os.environ["MBO_WG_VARIANT"] = "0"
dW2 = [torch.zeros_like(w) for w in Ws]
db2 = [torch.zeros_like(b) for b in bs]
pg(states, Wd, bd, a32, advs, logp_old, eps, entc, E, dW2, db2)
torch.cuda.synchronize()
print("deterministic:", all(torch.equal(a, b) for a, b in zip(dW + db, dW2 + db2)))

# timing
def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def torch_fb():
    for p in Ws + bs:
        p.grad = None
    _, _, _, _, _, loss = torch_loss()
    loss.backward()


t_native = timeit(lambda: pg(states, Wd, bd, a32, advs, logp_old, eps, entc, E, dW, db))
t_torch = timeit(torch_fb)
torch.backends.cuda.matmul.allow_tf32 = True
t_torch_tf32 = timeit(torch_fb)
flops = 3 * 2 * E * M * (F * H + 3 * H * H + H)
print("native %.3f ms (%.1f TFLOP/s)   torch fp32 %.3f ms   torch tf32 %.3f ms" % (t_native, flops / t_native / 1e9, t_torch, t_torch_tf32))
pg.check()
