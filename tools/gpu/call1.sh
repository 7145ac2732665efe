#!/bin/bash
# round-2 call 1: GPU tests on the new parity cases, bench baseline, ncu traffic of the fp16 MLP and the GP kernel
set -x
mkdir -p gpurun_out/r2c1
O=gpurun_out/r2c1
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > $O/smi.txt
nproc > $O/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:neural_af_tc -c 3 -o $O/mlp_fp16 python tools/tc_ncu.py fp16 1024 10000 > $O/ncu_mlp.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gp_posterior -c 3 -o $O/gp python tools/gp_ncu.py > $O/ncu_gp.log 2>&1
ncu -i $O/mlp_fp16.ncu-rep --page raw --csv > $O/mlp_fp16_raw.csv 2>/dev/null
ncu -i $O/gp.ncu-rep --page raw --csv > $O/gp_raw.csv 2>/dev/null
ls -la $O
