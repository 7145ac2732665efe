#!/bin/bash
O=gpurun_out/r2c17; mkdir -p $O
timeout 900 python -m pytest tests/test_dropin_gpu.py tests/test_comm_gpu.py tests/test_ppo_update_gpu.py tests/test_ppo_update_simt_gpu.py -m gpu -q > $O/tests.log 2>&1; tail -3 $O/tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-modes > $O/bench.json 2> $O/bench.err
python - <<'PY'
import json
b=json.loads(open("gpurun_out/r2c17/bench.json").read().splitlines()[-1])
print(b["ppo_rollout"]["env_steps_per_s"], [ (t["t_batch_s"], t["t_optim_s"]) for t in b["ppo_train"]])
PY
tail -3 $O/bench.err
