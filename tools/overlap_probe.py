"""Experiment: do the fp64 GP kernel and the tensor-core MLP kernel overlap when launched on two streams?
(needs the MLP kernel's shared memory small enough for a GP CTA to co-reside: build with -DMBO_TC_NSTAGE_TF32=4)"""
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
B, M, D, T, n = 512, 10000, 3, 30, 30
g = torch.Generator(device=dev); g.manual_seed(0)
cs = ops.CandidateSet(D, tail=torch.rand((M, D), generator=g, dtype=torch.float64, device=dev))
epi = torch.zeros(B, 4, device=dev); epi[:, 2] = 15; epi[:, 3] = 30
X = torch.rand((B, T, D), generator=g, dtype=torch.float64, device=dev)
Y = torch.randn((B, T), generator=g, dtype=torch.float64, device=dev)
st, _ = ops.gp_fit(torch.full((B,), n, dtype=torch.int32, device=dev), X, Y, torch.tensor([[0.716, 0.298, 0.186]] * B, dtype=torch.float64, device=dev),
                   torch.full((B,), 0.83, dtype=torch.float64, device=dev), torch.full((B,), 1e-6, dtype=torch.float64, device=dev))
mA = torch.randn((B, M), device=dev); sA = torch.rand((B, M), device=dev); lgA = torch.empty((B, M), device=dev)
mB = torch.empty((B, M), device=dev); sB = torch.empty((B, M), device=dev)
codes = [0, 1, 2, 3, 4, 12, 13]
s1, s2 = torch.cuda.Stream(), (torch.cuda.Stream(priority=-1) if os.environ.get('PROBE_PRIO', '1') == '1' else torch.cuda.Stream())
MODE = {"tf32": 1, "bf16x3": 2, "fp16": 3}[sys.argv[1] if len(sys.argv) > 1 else "fp16"]
def mlp(): ops.af_logits(pol, codes, cs, mA, sA, epi, B, mode=MODE, logits=lgA)
def gp(): ops.gp_posterior(st, T, cs, B, out_dtype=torch.float32, n_active_max=n, mean=mB, std=sB)
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return a.elapsed_time(e) / reps
def both():
    ev = torch.cuda.Event(); ev.record()
    s1.wait_event(ev); s2.wait_event(ev)
    with torch.cuda.stream(s2): mlp()      # persistent kernel first (high priority)
    with torch.cuda.stream(s1): gp()
    e1, e2 = torch.cuda.Event(), torch.cuda.Event()
    e1.record(s1); e2.record(s2)
    torch.cuda.current_stream().wait_event(e1); torch.cuda.current_stream().wait_event(e2)
def both_gp_first():
    ev = torch.cuda.Event(); ev.record()
    s1.wait_event(ev); s2.wait_event(ev)
    with torch.cuda.stream(s1): gp()
    with torch.cuda.stream(s2): mlp()
    e1, e2 = torch.cuda.Event(), torch.cuda.Event()
    e1.record(s1); e2.record(s2)
    torch.cuda.current_stream().wait_event(e1); torch.cuda.current_stream().wait_event(e2)
print("MBO_GP_TS=%s mode %d: mlp alone %.3f ms, gp alone %.3f ms, both (mlp first) %.3f ms, both (gp first) %.3f ms" %
      (os.environ.get("MBO_GP_TS"), MODE, timed(mlp), timed(gp), timed(both), timed(both_gp_first)))
pol.check()


def detail(first):
    """per-stream end times of one concurrent run (ms after the common start)"""
    torch.cuda.synchronize()
    ev = torch.cuda.Event(enable_timing=True); ev.record()
    s1.wait_event(ev); s2.wait_event(ev)
    em, eg = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    def run_mlp():
        with torch.cuda.stream(s2):
            mlp(); em.record(s2)
    def run_gp():
        with torch.cuda.stream(s1):
            gp(); eg.record(s1)
    (run_mlp(), run_gp()) if first == "mlp" else (run_gp(), run_mlp())
    torch.cuda.synchronize()
    return ev.elapsed_time(em), ev.elapsed_time(eg)


for first in ("mlp", "gp"):
    for _ in range(2):
        a, b_ = detail(first)
    print("%s launched first: mlp ends at %.3f ms, gp ends at %.3f ms" % (first, a, b_))
