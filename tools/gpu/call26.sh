#!/bin/bash
O=gpurun_out/r2c26; mkdir -p $O
python - <<'PY' > $O/train.log 2>&1
import os, sys, torch, json
sys.path.insert(0, os.getcwd())
import bench
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
r = bench.ppo_train_rate(dev, 1, 0, None, 4, "C3 reference batch")
print("RESULT", json.dumps(r))
r = bench.ppo_train_rate(dev, 1, 0, 1024, 1, "C3 1024 episodes per GPU")
print("RESULT", json.dumps(r))
PY
grep RESULT $O/train.log || tail -20 $O/train.log
