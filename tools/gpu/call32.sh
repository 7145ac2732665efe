#!/bin/bash
O=gpurun_out/r2c32; mkdir -p $O
timeout 900 python -m pytest tests/test_af_gpu.py tests/test_golden_round_gpu.py -m gpu -q -x > $O/tests.log 2>&1; tail -4 $O/tests.log
python bench.py --steps 30 --warmup 5 --no-modes --no-cpu-baseline --no-ppo > $O/bench.json 2> $O/bench.err
python -c "
import json
d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1])
print('ms_per_step', d['ms_per_step'], 'mlp ms', d['roofline']['avg_launch_ms'], 'frac', d['roofline']['frac'], 'gp', d['roofline']['gp_posterior']['avg_launch_ms'])
"
python tools/tc_trace.py fp16 > $O/trace.log 2>&1; sed -n 1,6p $O/trace.log; sed -n 10,25p $O/trace.log
