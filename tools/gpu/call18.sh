#!/bin/bash
O=gpurun_out/r2c18; mkdir -p $O
MBO_NO_GRAPHS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/rollout_launches.csv python tools/rollout_ncu.py 1 > $O/ncu.log 2>&1
python tools/summarize_launches.py $O/rollout_launches.csv | head -40
