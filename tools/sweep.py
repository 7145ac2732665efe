"""BASELINE configs[4]: synthetic sweep over B episodes x M candidates x n observations x D dims (one GPU).
Per point: GP posterior (fp64) and NeuralAF MLP (4x200, N(0,0.01) weights of width F = D + 4) timed separately with
CUDA events (3 repetitions after one warm-up), shared candidate table unless --per-episode (fp32 [B,M,D] streaming).
Prints a markdown table: ms and achieved algorithmic TFLOP/s (GP: n^2 + n(3D+8) FLOP, MLP: 2*sum(in*out)) and GB/s."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metabo_b200 import ops, _lib as L
dev = torch.device("cuda:0")
per_episode = "--per-episode" in sys.argv
g = torch.Generator(device=dev); g.manual_seed(0)


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


def policy(F):
    rng = np.random.RandomState(0)
    dims = [F, 200, 200, 200, 200, 1]
    W = [torch.as_tensor(rng.normal(0, 0.01, (dims[i + 1], dims[i])).astype(np.float32), device=dev) for i in range(5)]
    b = [torch.as_tensor(rng.normal(0, 0.01, dims[i + 1]).astype(np.float32), device=dev) for i in range(5)]
    h = ops.PolicyHandle(); h.set_weights(W, b, "relu")
    return h


print("| B | M | n | D | cands | GP fp64 ms | GP TF/s (alg) | MLP tf32 ms | MLP TF/s | MLP bf16x3 ms | cand MB | total G AF evals/s (tf32) |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
grid = [(1, 1024, 10, 2), (8, 4096, 30, 2), (64, 16384, 30, 3), (512, 4096, 30, 3), (1024, 16384, 50, 3), (1024, 16384, 100, 3),
        (4096, 4096, 30, 4), (8192, 1024, 10, 2), (8192, 16384, 30, 3), (2048, 65536, 30, 3), (1024, 65536, 100, 3)]
if per_episode:
    grid = [(1024, 16384, 30, 3), (4096, 16384, 30, 3), (8192, 16384, 10, 2)]
for B, M, n, D in grid:
    T = n
    F = D + 4
    pol = policy(F)
    codes = [L.FEAT_MEAN, L.FEAT_STD] + [L.FEAT_X0 + d for d in range(D)] + [L.FEAT_TIMESTEP, L.FEAT_BUDGET]
    if per_episode:
        cs = ops.CandidateSet(D, head=torch.rand((B, M, D), generator=g, dtype=torch.float32, device=dev))
    else:
        cs = ops.CandidateSet(D, tail=torch.rand((M, D), generator=g, dtype=torch.float64, device=dev))
    X = torch.rand((B, T, D), generator=g, dtype=torch.float64, device=dev)
    Y = torch.randn((B, T), generator=g, dtype=torch.float64, device=dev)
    st, status = ops.gp_fit(torch.full((B,), n, dtype=torch.int32, device=dev), X, Y,
                            torch.full((B, D), 0.3, dtype=torch.float64, device=dev),
                            torch.ones(B, dtype=torch.float64, device=dev), torch.full((B,), 1e-4, dtype=torch.float64, device=dev))
    mean = torch.empty((B, M), device=dev); std = torch.empty((B, M), device=dev); lg = torch.empty((B, M), device=dev)
    epi = torch.zeros(B, 4, device=dev); epi[:, 2] = n; epi[:, 3] = 30
    tg = timeit(lambda: ops.gp_posterior(st, T, cs, B, out_dtype=torch.float32, n_active_max=n, mean=mean, std=std))
    t1 = timeit(lambda: ops.af_logits(pol, codes, cs, mean, std, epi, B, mode=L.MLP_TF32, logits=lg))
    t2 = timeit(lambda: ops.af_logits(pol, codes, cs, mean, std, epi, B, mode=L.MLP_BF16X3, logits=lg))
    pol.check()
    c = B * M
    gpf = c * (n * n + n * (3 * D + 8)) / tg / 1e9
    mlf = c * 2 * (F * 200 + 3 * 200 * 200 + 200) / t1 / 1e9
    mb = (B * M * D * 4 / 1e6) if per_episode else (M * D * 8 / 1e6)
    print("| %d | %d | %d | %d | %.2e | %.3f | %.2f | %.3f | %.0f | %.3f | %.1f | %.2f |" %
          (B, M, n, D, c, tg, gpf, t1, mlf, t2, mb, c / (tg + t1) / 1e6), flush=True)
    del st, mean, std, lg, cs
    torch.cuda.empty_cache()
