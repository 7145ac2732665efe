#!/bin/bash
O=gpurun_out/r2c30; mkdir -p $O
for dbg in 0 2 4 1; do
MBO_TC_DEBUG_NOLOAD=$dbg timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/dbg_$dbg.csv python tools/upd_once.py > $O/ncu_$dbg.log 2>&1
echo "== MBO_TC_DEBUG_NOLOAD=$dbg"; python tools/summarize_launches.py $O/dbg_$dbg.csv | grep neural_af
done
