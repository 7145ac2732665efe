#!/bin/bash
# round-2 final single-GPU measurements (second pass, after the selection-epilogue default and the update rework):
# tests, smoke, bench (both arms), ncu launch lists + full captures, configs, sweep
set -x
O=gpurun_out/r2fin3; mkdir -p $O
nproc > $O/nproc.txt; lscpu | grep "Model name" >> $O/nproc.txt
timeout 1200 python -m pytest tests -m gpu -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-modes --no-ppo > $O/ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:neural_af_tc -c 3 -o $O/mlp_fp16 python tools/tc_ncu.py fp16 1024 10000 topk > $O/ncu_mlp.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gp_posterior -c 3 -o $O/gp python tools/gp_ncu.py > $O/ncu_gp.log 2>&1
ncu -i $O/mlp_fp16.ncu-rep --page raw --csv > $O/mlp_fp16_raw.csv 2>/dev/null
ncu -i $O/gp.ncu-rep --page raw --csv > $O/gp_raw.csv 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/upd_launches.csv python tools/upd_once.py > $O/ncu_upd.log 2>&1
timeout 900 python tools/run_configs.py C1 C2 C3 C4 > $O/configs.jsonl 2> $O/configs.err
timeout 300 python tools/sweep.py > $O/sweep_1gpu.md 2> $O/sweep.err
timeout 600 python tools/ppo_learning_curve.py > $O/learning_curve.json 2> $O/learning_curve.err
tail -3 $O/tests.log; cat $O/smoke.log | tail -2; tail -2 $O/bench.err; tail -3 $O/configs.err
