#!/bin/bash
O=gpurun_out/r2c36; mkdir -p $O
python tools/rollout_ncu.py 4 40 > $O/rollout40_graphs.log 2>&1; tail -2 $O/rollout40_graphs.log
MBO_NO_GRAPHS=1 python tools/rollout_ncu.py 3 40 > $O/rollout40_eager.log 2>&1; tail -1 $O/rollout40_eager.log
MBO_NO_GRAPHS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/rollout40.csv python tools/rollout_ncu.py 1 40 > $O/ncu.log 2>&1
python tools/summarize_launches.py $O/rollout40.csv > $O/rollout40.md; head -45 $O/rollout40.md
