"""Pipeline timeline of one CTA of the forward GEMM of the PPO update (debug build: tools/build_trace_lib.sh, then
MBO_LIB_PATH=metabo_b200/libmetabo_b200_trace.so python tools/upd_trace.py): clock64 stamps of the loader thread 0, the MMA
thread, the weight-stage thread and the epilogue, printed as cycles relative to the CTA's first stamp."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from metabo_b200 import ops, _lib as L
E, M, H, F = 60, 966, 200, 6
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
dims = [F, H, H, H, H, 1]
W = [(torch.randn(dims[i + 1], dims[i], generator=g) * 0.08).to(dev) for i in range(5)]
b = [(torch.randn(dims[i + 1], generator=g) * 0.05).to(dev) for i in range(5)]
st = torch.randn(E, M, F, generator=g).to(dev)
ac = torch.randint(0, M, (E,), generator=g).to(dev).to(torch.int32)
pg = ops.PPOPolicyGrad(E, M, F, list(range(F)), H, dev)
for _ in range(3):
    pg.logp_entropy(st, W, b, ac)          # forward only: the last forward GEMM's stamps stay in the trace buffer
torch.cuda.synchronize()
buf = (C.c_longlong * 4096)()
f = L.lib().mbo_upd_trace_read
f.argtypes = [C.POINTER(C.c_longlong)]
assert f(buf) == 0
t = np.array(buf[:], dtype=np.int64).reshape(8, 512)
t0 = min(int(x) for x in t[:4].reshape(-1) if x > 0)
rel = lambda x: int(x) - t0
NKT = 26
print("loader thread 0 (even K tiles), per own K tile: [iteration start, own copies landed, operand stage free, chunks split, "
      "fence.proxy.async done, arrived]")
for kt in range(0, NKT, 2):
    print("  kt %2d: %6d %6d %6d %6d %6d %6d" % ((kt,) + tuple(rel(t[0, 4 * kt + i]) for i in range(3)) +
                                          (rel(t[4, 2 * kt]), rel(t[4, 2 * kt + 1]), rel(t[0, 4 * kt + 3]))))
print("MMA thread, per K tile: [wait start, stage full, MMAs + commit issued]")
for kt in range(NKT):
    print("  kt %2d: %6d %6d %6d" % ((kt,) + tuple(rel(t[1, 3 * kt + i]) for i in range(3))))
print("weight thread, per K tile: [wait start, bulk copies issued]")
for kt in range(NKT):
    print("  kt %2d: %6d %6d" % ((kt,) + tuple(rel(t[2, 2 * kt + i]) for i in range(2))))
print("epilogue (thread 0): wait start %d, accumulator ready %d, stores issued %d" % tuple(rel(t[3, i]) for i in range(3)))
