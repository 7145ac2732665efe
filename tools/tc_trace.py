# Pipeline trace of the tcgen05 NeuralAF kernel: run with the -DMBO_TC_TRACE library (tools/build_trace_lib.sh) and
# print the clock64 stamps of CTA 0, tile 2 (cycles relative to the MMA warp's "features ready").
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["MBO_LIB_PATH"] = os.path.join(ROOT, "metabo_b200", "libmetabo_b200_trace.so")
sys.path.insert(0, ROOT)
import numpy as np, torch
from metabo_b200 import ops, _lib as L
from oracle import metabo_oracle as O
mode_name = sys.argv[1] if len(sys.argv) > 1 else "fp16"
mode = L.MLP_MODES[mode_name]
dev = torch.device("cuda:0")
pw = O.PolicyWeights.from_npz(os.path.join(ROOT, "tests/golden/weights_hm3_1988.npz"))
layers = pw.layers("policy_net")
pol = ops.PolicyHandle(); pol.set_weights([w.cuda() for w, _ in layers], [b.cuda() for _, b in layers], "relu")
codes = [0, 1, 2, 3, 4, 12, 13]
rng = np.random.RandomState(0)
B, M = 1024, 10000
cs = ops.CandidateSet(3, tail=torch.as_tensor(rng.rand(M, 3), device=dev))
mean = torch.randn(B, M, device=dev); std = torch.rand(B, M, device=dev)
epi = torch.zeros(B, 4, device=dev); epi[:, 2] = 15; epi[:, 3] = 30
for _ in range(2):
    ops.af_logits(pol, codes, cs, mean, std, epi, B, mode=mode)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 8192)()
lib = L.lib()
lib.mbo_debug_tc_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert lib.mbo_debug_tc_trace(buf, 8192) == 0
raw = np.array(buf[:], dtype=np.int64)
trs = {0: raw[:2048].reshape(4, 512), 82: raw[2048:4096].reshape(4, 512)}
cta = raw[4096:4096 + 148 * 3].reshape(148, 3)
smid = cta[:, 2] >> 20
cta[:, 2] &= (1 << 20) - 1
print("per-CTA: cycles min/med/max", cta[:, 0].min(), int(np.median(cta[:, 0])), cta[:, 0].max(), " ns min/med/max", cta[:, 1].min(), int(np.median(cta[:, 1])), cta[:, 1].max(), " tiles", cta[:, 2].min(), cta[:, 2].max(), " GHz %.3f" % (cta[:, 0].sum() / cta[:, 1].sum()))
st = raw[5120:5120 + 148 * 8].reshape(148, 8)
order = np.argsort(cta[:, 0])
print("per-CTA wait cycles per tile (fastest, median, slowest CTA): total | MMA waits feat, a_ready, w_full | epi warp0 d_full | feature build+gather")
for i in (order[0], order[74], order[-1]):
    n = cta[i, 2]
    print("  CTA %3d: %6.0f | %6.0f %6.0f %6.0f | %6.0f | %6.0f" % (i, cta[i, 0] / n, st[i, 0] / n, st[i, 1] / n, st[i, 2] / n, st[i, 3] / n, st[i, 4] / n))
print('cycles/tile by smid (sorted by smid):')
idx = np.argsort(smid)
print(' '.join('%d:%d' % (smid[i], cta[i, 0] // cta[i, 2]) for i in idx))
print("CTA 0: cycles", cta[0, 0], "ns", cta[0, 1], " cycles/tile %.0f" % (cta[0, 0] / cta[0, 2]))
for cta_id, tr in trs.items():
    print('######## CTA', cta_id)
    t0 = tr[0, 0]
    def rel(tile, slot):
        v = tr[tile, slot]
        return int(v - t0) if v else None
    nch = {"tf32": 5, "bf16x3": 7, "fp16": 7}[mode_name]
    for tile in (0, 1):
        print("== traced tile %d (%s) ==" % (tile, mode_name))
        print("MMA: feat ready", rel(tile, 0), " L0 issued", rel(tile, 1))
        for l in (1, 2, 3):
            print("MMA L%d: a_ready seen" % l, [rel(tile, 10 * l + c) for c in range(nch)], " committed", rel(tile, 10 * l + 8))
        print("FEAT warps: s_free seen", rel(tile, 300), " E0 chunk arrives", [rel(tile, 310 + i) for i in range(7)], " build done", rel(tile, 302), " gather done", rel(tile, 301))
        for half in (0, 1):
            base = 100 + 100 * half
            for l in (0, 1, 2):
                print("EPI h%d E%d: d_full" % (half, l), rel(tile, base + 20 * l), " arrives", [rel(tile, base + 20 * l + 1 + i) for i in range((nch + 1) // 2)])
            print("EPI h%d E3: d_full" % half, rel(tile, base + 70), " summed", rel(tile, base + 71), " stored", rel(tile, base + 72),
                  " features", rel(tile, base + 80), rel(tile, base + 81))
