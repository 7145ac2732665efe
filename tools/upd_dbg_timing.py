"""Timing experiments on the SS-form update GEMMs: MBO_UPD_DBG bit mask (1 no epilogue, 2 no MMAs, 4 no A loads, 8 no weight
loads); results are garbage, only the time of the whole call is of interest.  The hooks live in upd_gemm_kernel (run with
MBO_UPD_FWD_TS=0 to time the SS forward; bit 4 applies to its dgrad instantiation only) and in upd_dgrad_persist_kernel; the
TS-form kernels (defaults for the forward) have none.  Tables of the round: profiles/r01c_update_launches.md."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from metabo_b200 import ops
E, M, H, F = 60, 966, 200, 6
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
dims = [F, H, H, H, H, 1]
W = [(torch.randn(dims[i + 1], dims[i], generator=g) * 0.08).to(dev) for i in range(5)]
b = [(torch.randn(dims[i + 1], generator=g) * 0.05).to(dev) for i in range(5)]
st = torch.randn(E, M, F, generator=g).to(dev)
ac = torch.randint(0, M, (E,), generator=g).to(dev).to(torch.int32)
adv = torch.randn(E, generator=g).to(dev)
lpo = torch.full((E,), -6.9, device=dev)
pg = ops.PPOPolicyGrad(E, M, F, list(range(F)), H, dev)
dW = [torch.zeros_like(w) for w in W]; db = [torch.zeros_like(x) for x in b]
def run(fwd_only):
    if fwd_only:
        pg.logp_entropy(st, W, b, ac)
    else:
        pg(st, W, b, ac, adv, lpo, 0.15, 0.01, E, dW, db)
for dbg in (0, 1, 2, 4, 8, 3, 6, 12, 14, 15):
    os.environ["MBO_UPD_DBG"] = str(dbg)
    out = []
    for fwd_only in (True, False):
        for _ in range(3):
            run(fwd_only)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            run(fwd_only)
        e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / 20 * 1e3)
    print("dbg %2d: forward-only call %7.1f us   full call %7.1f us" % (dbg, out[0], out[1]), flush=True)
pg.ws[0] = 0
