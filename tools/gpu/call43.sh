#!/bin/bash
O=gpurun_out/r2c43; mkdir -p $O
MBO_NO_GRAPHS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file $O/c4.csv python tools/c4_rollout_ncu.py 1000 > $O/ncu.log 2>&1
python tools/summarize_launches.py $O/c4.csv > $O/c4.md; head -24 $O/c4.md; tail -3 $O/ncu.log
