#!/bin/bash
set -x
O=gpurun_out/r2c6; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
grep -n "^FAILED\|^ERROR\|passed\|failed" $O/tests.log | tail -20
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-modes --no-ppo > $O/bench_mma.json 2> $O/bench_mma.err
MBO_GP_SIMT=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-modes --no-ppo > $O/bench_simt.json 2> $O/bench_simt.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gp_posterior -c 3 -o $O/gp_mma python tools/gp_ncu.py > $O/ncu_gp.log 2>&1
ncu -i $O/gp_mma.ncu-rep --page raw --csv > $O/gp_mma_raw.csv 2>/dev/null
timeout 300 python tools/sweep.py > $O/sweep.log 2>&1
python - <<'PY'
import json
for f in ("bench_mma","bench_simt"):
    try:
        b=json.loads(open("gpurun_out/r2c6/%s.json"%f).read().splitlines()[-1])
        print(f, b["ms_per_step"], b["value"], b["roofline"]["gp_posterior"], b["roofline"]["avg_launch_ms"])
    except Exception as e: print(f, "ERR", e)
PY
tail -30 $O/sweep.log
