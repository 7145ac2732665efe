#!/bin/bash
O=gpurun_out/r2c34; mkdir -p $O
for dbg in 0 1 16 2 8 10 27; do
MBO_UPD_DBG=$dbg timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/dbg_$dbg.csv python tools/upd_once.py > $O/ncu_$dbg.log 2>&1
echo "== MBO_UPD_DBG=$dbg"; python tools/summarize_launches.py $O/dbg_$dbg.csv 2>/dev/null | grep chain
done
