"""Soak test of the tcgen05 NeuralAF kernel: random batch / candidate-set / feature shapes in all tensor-core modes,
every result compared with the fp32 CUDA-core kernel, the error word checked after every launch, the same launch
repeated to catch run-to-run differences (races).  python tools/tc_soak.py [iterations] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from metabo_b200 import ops, _lib as L
from oracle import metabo_oracle as O
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 150
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = torch.device("cuda:0")
rng = np.random.RandomState(seed)
pw = O.PolicyWeights.from_npz("tests/golden/weights_hm3_1988.npz")
layers = pw.layers("policy_net")
pol7 = ops.PolicyHandle(); pol7.set_weights([w.cuda() for w, _ in layers], [b.cuda() for _, b in layers], "relu")
worst = {}
t0 = time.time()
for it in range(iters):
    D = int(rng.choice([2, 3, 3, 3, 4, 6]))
    if D == 3:
        pol, codes = pol7, [0, 1, 2, 3, 4, 12, 13]
    else:
        F = D + 4
        dims = [F, 200, 200, 200, 200, 1]
        W = [torch.as_tensor(rng.normal(0, 0.08, (dims[i + 1], dims[i])).astype(np.float32), device=dev) for i in range(5)]
        b = [torch.as_tensor(rng.normal(0, 0.05, dims[i + 1]).astype(np.float32), device=dev) for i in range(5)]
        pol = ops.PolicyHandle(); pol.set_weights(W, b, "relu")
        codes = [L.FEAT_MEAN, L.FEAT_STD] + [L.FEAT_X0 + d for d in range(D)] + [L.FEAT_TIMESTEP, L.FEAT_BUDGET]
    B = int(rng.choice([1, 2, 3, 7, 16, 64, 150, 300]))
    M = int(rng.choice([1, 5, 127, 128, 129, 961, 1733, 5000, 10000]))
    kind = int(rng.randint(3))
    if kind == 0:
        cs = ops.CandidateSet(D, tail=torch.as_tensor(rng.rand(M, D), device=dev))
    elif kind == 1 and M >= 10:
        k = 5
        nls = M // k
        M = nls * k
        cubes = ops.cubes_around(torch.as_tensor(rng.rand(B, k, D), device=dev), 0.1)
        cs = ops.CandidateSet(D, tail=torch.as_tensor(rng.rand(nls, D), device=dev), cubes=cubes)
    else:
        nh = min(5, M)
        cs = ops.CandidateSet(D, head=torch.as_tensor(rng.rand(B, nh, D), device=dev),
                              tail=torch.as_tensor(rng.rand(max(M - nh, 1), D), device=dev))
        M = cs.M
    mean = torch.as_tensor(rng.randn(B, M).astype(np.float32), device=dev)
    std = torch.as_tensor(rng.rand(B, M).astype(np.float32), device=dev)
    epi = torch.as_tensor(np.stack([rng.randn(B), rng.rand(B), rng.randint(0, 30, B).astype(float), np.full(B, 30.)], 1).astype(np.float32), device=dev)
    only = int(os.environ.get("SOAK_ONLY", -1))
    if only >= 0 and it != only:
        continue
    ref = ops.af_logits(pol, codes, cs, mean, std, epi, B, mode=L.MLP_FP32)
    scale = float(ref.abs().max()) + 1e-30
    # a random net's logits can cancel to ~0 (max |logit| 0.04 with O(1) hidden activations): the 11-bit modes are then
    # judged on the scale of the head layer's terms, not of their cancelled sum
    w_head = pol._keep[0][-1]
    scale11 = max(scale, 0.02 * float(w_head.abs().sum()))
    for name, mode, tol in (("fp16", L.MLP_FP16, 2e-2), ("tf32", L.MLP_TF32, 2e-2), ("bf16x3", L.MLP_BF16X3, 1e-3)):
        got = ops.af_logits(pol, codes, cs, mean, std, epi, B, mode=mode)
        got2 = ops.af_logits(pol, codes, cs, mean, std, epi, B, mode=mode)
        torch.cuda.synchronize(); pol.check()
        err = float((got - ref).abs().max()) / (scale if name == "bf16x3" else scale11)
        if only >= 0:
            d = (got - ref).abs()
            idx = int(d.argmax())
            print(name, "err %.3e" % err, "scale %.3e" % scale, "B M D kind", B, M, D, kind, "argmax row/col", idx // M, idx % M, "n bad(>1e-2)", int((d > 1e-2 * scale).sum()), flush=True)
            continue
        worst[name] = max(worst.get(name, 0.0), err)
        assert torch.equal(got, got2), ("run-to-run difference", it, name, B, M, D, kind)
        assert err <= tol, ("mismatch", it, name, B, M, D, kind, err)
print("soak ok: %d iterations, %.1f s, worst relative errors %s" % (iters, time.time() - t0, {k: "%.2e" % v for k, v in worst.items()}))
