#!/bin/bash
O=gpurun_out/r2c35; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:neural_af_tc -c 4 -o $O/k2store python tools/upd_once.py > $O/ncu.log 2>&1
ncu -i $O/k2store.ncu-rep --page raw --csv > $O/k2store_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:upd_dgrad_chain -c 2 -o $O/chain python tools/upd_once.py > $O/ncu2.log 2>&1
ncu -i $O/chain.ncu-rep --page raw --csv > $O/chain_raw.csv 2>/dev/null
ls -la $O
