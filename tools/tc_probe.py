# quick probe: TC logits vs fp32 SIMT logits on random states, both modes, several sizes
import sys, numpy as np, torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from metabo_b200 import ops, _lib as L
from oracle import metabo_oracle as O
dev = torch.device("cuda:0")
pw = O.PolicyWeights.from_npz("tests/golden/weights_hm3_1988.npz")
layers = pw.layers("policy_net")
pol = ops.PolicyHandle(); pol.set_weights([w.cuda() for w,_ in layers],[b.cuda() for _,b in layers],"relu")
codes = [0,1,2,3,4,12,13]
rng = np.random.RandomState(0)
for B, M in [(1,128),(1,100),(3,1733),(64,10000)]:
    x = torch.as_tensor(rng.rand(M,3), device=dev)
    cs = ops.CandidateSet(3, tail=x)
    mean = torch.as_tensor(rng.randn(B,M).astype(np.float32), device=dev)
    std = torch.as_tensor(rng.rand(B,M).astype(np.float32), device=dev)
    epi = torch.as_tensor(np.stack([np.zeros(B), np.full(B,0.4), rng.randint(0,30,B).astype(float), np.full(B,30.)],1).astype(np.float32), device=dev)
    ref = ops.af_logits(pol, codes, cs, mean, std, epi, B, mode=0)
    for name, mode in (("fp16",3),("tf32",1),("bf16x3",2)):
        try:
            got = ops.af_logits(pol, codes, cs, mean, std, epi, B, mode=mode)
            torch.cuda.synchronize(); pol.check()
            err = (got-ref).abs().max().item()/ref.abs().max().item()
            print(B, M, name, "max rel err %.3e"%err, "ref max %.2f"%ref.abs().max().item(), "nan", torch.isnan(got).sum().item(), flush=True)
        except Exception as e:
            print(B, M, name, "FAILED", e, flush=True)
# timing
B, M = 1024, 10000
x = torch.as_tensor(rng.rand(M,3), device=dev); cs = ops.CandidateSet(3, tail=x)
mean = torch.randn(B,M,device=dev); std = torch.rand(B,M,device=dev)
epi = torch.zeros(B,4,device=dev); epi[:,2]=15; epi[:,3]=30
for name, mode in (("fp16",3),("tf32",1),("bf16x3",2),("fp32",0)):
    try:
        lg = ops.af_logits(pol, codes, cs, mean, std, epi, B, mode=mode); torch.cuda.synchronize()
        s,e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(3): ops.af_logits(pol, codes, cs, mean, std, epi, B, mode=mode, logits=lg)
        e.record(); torch.cuda.synchronize(); pol.check()
        ms = s.elapsed_time(e)/3
        print(name, "%.3f ms"%ms, "%.1f TFLOP/s"%(B*M*243200/ms/1e9), flush=True)
    except Exception as ex:
        print(name, "FAILED", ex, flush=True)
