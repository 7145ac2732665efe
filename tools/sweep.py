"""BASELINE configs[4]: synthetic sweep over B episodes x M candidates x n observations x D dims, on 1 GPU or under
torchrun on N GPUs (every rank evaluates its own B episodes: episodes are independent units, no collective on the data
path; times are the max over ranks, rates the sum).
Per point: GP posterior (fp64) and NeuralAF MLP (4x200, N(0,0.01) weights of width F = D + 4) in fp16 (the mode bench.py
times), tf32 and bf16x3, timed separately with CUDA events (3 repetitions after one warm-up), shared candidate table
unless --per-episode (fp32 [B,M,D] streaming).
Prints a markdown table: ms and achieved algorithmic TFLOP/s (GP: n^2 + n(3D+8) FLOP, MLP: 2*sum(in*out)) per GPU."""
import os, sys
import numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metabo_b200 import ops, _lib as L
rank, world, lrank = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lrank)
dev = torch.device("cuda", lrank)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
per_episode = "--per-episode" in sys.argv
g = torch.Generator(device=dev); g.manual_seed(rank)


def rmax(x):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record(); torch.cuda.synchronize()
    return rmax(s.elapsed_time(e) / reps)


def policy(F):
    rng = np.random.RandomState(0)
    dims = [F, 200, 200, 200, 200, 1]
    W = [torch.as_tensor(rng.normal(0, 0.01, (dims[i + 1], dims[i])).astype(np.float32), device=dev) for i in range(5)]
    b = [torch.as_tensor(rng.normal(0, 0.01, dims[i + 1]).astype(np.float32), device=dev) for i in range(5)]
    h = ops.PolicyHandle(); h.set_weights(W, b, "relu")
    return h


if rank == 0:
    print("%d GPU(s); B = episodes PER GPU; TFLOP/s per GPU; last column = all GPUs, GP + fp16 MLP\n" % world)
    print("| B | M | n | D | cands/GPU | GP fp64 ms | GP TF/s (alg) | MLP fp16 ms | MLP fp16 TF/s | MLP tf32 ms | MLP bf16x3 ms | cand MB | total G AF evals/s |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
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
    t0 = timeit(lambda: ops.af_logits(pol, codes, cs, mean, std, epi, B, mode=L.MLP_FP16, logits=lg))
    t1 = timeit(lambda: ops.af_logits(pol, codes, cs, mean, std, epi, B, mode=L.MLP_TF32, logits=lg))
    t2 = timeit(lambda: ops.af_logits(pol, codes, cs, mean, std, epi, B, mode=L.MLP_BF16X3, logits=lg))
    pol.check()
    c = B * M
    gpf = c * (n * n + n * (3 * D + 8)) / tg / 1e9
    mlf = c * 2 * (F * 200 + 3 * 200 * 200 + 200) / t0 / 1e9
    mb = (B * M * D * 4 / 1e6) if per_episode else (M * D * 8 / 1e6)
    if rank == 0:
        print("| %d | %d | %d | %d | %.2e | %.3f | %.2f | %.3f | %.0f | %.3f | %.3f | %.1f | %.2f |" %
              (B, M, n, D, c, tg, gpf, t0, mlf, t1, t2, mb, world * c / (tg + t0) / 1e6), flush=True)
    del st, mean, std, lg, cs
    torch.cuda.empty_cache()
if world > 1:
    dist.destroy_process_group()
