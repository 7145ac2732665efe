#!/bin/bash
O=gpurun_out/r2c15; mkdir -p $O
timeout 300 python tools/tc_probe.py > $O/tc_probe.log 2>&1; tail -5 $O/tc_probe.log
timeout 600 python -m pytest tests/test_af_gpu.py tests/test_golden_round_gpu.py -m gpu -q -x > $O/tests.log 2>&1; tail -3 $O/tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-modes --no-ppo > $O/bench.json 2> $O/bench.err
python - <<'PY'
import json
b=json.loads(open("gpurun_out/r2c15/bench.json").read().splitlines()[-1])
print(b["ms_per_step"], b["value"], b["roofline"]["frac"], b["roofline"]["avg_launch_ms"], b["roofline"]["gp_posterior"])
PY
