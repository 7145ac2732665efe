#!/bin/bash
O=gpurun_out/r2c33; mkdir -p $O
timeout 1200 python -m pytest tests/test_ppo_update_gpu.py tests/test_dropin_gpu.py tests/test_comm_gpu.py -m gpu -q -x > $O/tests.log 2>&1; tail -5 $O/tests.log
python - <<'PY' > $O/upd.log 2>&1
import os, sys, torch, json
sys.path.insert(0, os.getcwd())
import bench
dev = torch.device("cuda:0")
for ch in ("1", "0"):
    os.environ["MBO_UPD_DGRAD_CHAIN"] = ch
    r = bench.ppo_update_rate(dev)
    print("MBO_UPD_DGRAD_CHAIN=%s policy_grad %.4f ms (%.1f TFLOP/s alg)" % (ch, r["policy_grad_ms"], r["algorithmic_tflops"]))
PY
tail -3 $O/upd.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/upd.csv python tools/upd_once.py > $O/ncu.log 2>&1
python tools/summarize_launches.py $O/upd.csv | head -14
