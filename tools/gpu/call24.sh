#!/bin/bash
O=gpurun_out/r2c24; mkdir -p $O
timeout 900 python -m pytest tests/test_ppo_update_gpu.py -m gpu -q > $O/tests.log 2>&1; tail -15 $O/tests.log
python - <<'PY' > $O/upd.log 2>&1
import os, sys, torch
sys.path.insert(0, os.getcwd())
import bench
dev = torch.device("cuda:0")
for k2 in ("1", "0"):
    os.environ["MBO_UPD_FWD_K2"] = k2
    r = bench.ppo_update_rate(dev)
    print("MBO_UPD_FWD_K2=%s policy_grad %.4f ms (%.1f TFLOP/s alg), torch autograd %.3f ms" % (k2, r["policy_grad_ms"], r["algorithmic_tflops"], r["torch_fp32_autograd_ms"]))
PY
cat $O/upd.log | tail -4
sed 's/r2c23/r2c24/' tools/gpu/call23.sh > /tmp/c23.sh; sed -i 's/for k2 in 1 0/for k2 in 1/' /tmp/c23.sh; bash /tmp/c23.sh
