#!/bin/bash
O=gpurun_out/r2c25; mkdir -p $O
timeout 1200 python -m pytest tests/test_ppo_update_gpu.py tests/test_dropin_gpu.py tests/test_comm_gpu.py -m gpu -q -x > $O/tests.log 2>&1; tail -5 $O/tests.log
python - <<'PY' > $O/upd.log 2>&1
import os, sys, torch, json
sys.path.insert(0, os.getcwd())
import bench
dev = torch.device("cuda:0")
for fk in ("1", "0"):
    os.environ["MBO_UPD_FORK"] = fk
    r = bench.ppo_update_rate(dev)
    print("MBO_UPD_FORK=%s policy_grad %.4f ms (%.1f TFLOP/s alg)" % (fk, r["policy_grad_ms"], r["algorithmic_tflops"]))
os.environ["MBO_UPD_FORK"] = "1"
r = bench.ppo_train_rate(dev, 1, 0, 40, 4, "C3 reference batch")
print(json.dumps(r))
PY
tail -4 $O/upd.log
