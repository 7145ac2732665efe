#!/bin/bash
O=gpurun_out/r2c38; mkdir -p $O
timeout 1200 python -m pytest tests/test_dropin_gpu.py tests/test_golden_round_gpu.py tests/test_comm_gpu.py -m gpu -q -x > $O/tests.log 2>&1; tail -3 $O/tests.log
MBO_NO_GRAPHS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/rollout40.csv python tools/rollout_ncu.py 1 40 > $O/ncu.log 2>&1
python tools/summarize_launches.py $O/rollout40.csv > $O/rollout40.md; head -40 $O/rollout40.md
