#!/bin/bash
set -x
O=gpurun_out/r2c8; mkdir -p $O
timeout 600 python -m pytest tests/test_gp_gpu.py tests/test_golden_round_gpu.py tests/test_af_gpu.py -m gpu -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
tail -5 $O/tests.log
timeout 300 python tools/gp_probe.py > $O/gp_probe_mma.log 2>&1
MBO_GP_SIMT=1 timeout 300 python tools/gp_probe.py > $O/gp_probe_simt.log 2>&1
cat $O/gp_probe_mma.log $O/gp_probe_simt.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-modes --no-ppo > $O/bench_mma.json 2> $O/bench_mma.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gp_posterior -c 3 -o $O/gp_mma python tools/gp_ncu.py > $O/ncu_gp.log 2>&1
ncu -i $O/gp_mma.ncu-rep --page raw --csv > $O/gp_mma_raw.csv 2>/dev/null
python - <<'PY'
import json
b=json.loads(open("gpurun_out/r2c8/bench_mma.json").read().splitlines()[-1])
print(b["ms_per_step"], b["value"], b["roofline"]["gp_posterior"], b["roofline"]["avg_launch_ms"])
PY
