#!/bin/bash
O=gpurun_out/r2c40; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/tests.log 2>&1; tail -3 $O/tests.log
python - <<'PY' > $O/train.log 2>&1
import os, sys, torch, json
sys.path.insert(0, os.getcwd())
import bench
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
r = bench.ppo_train_rate(dev, 1, 0, None, 4, "C3 reference batch")
print("RESULT", json.dumps(r))
PY
grep RESULT $O/train.log || tail -20 $O/train.log
