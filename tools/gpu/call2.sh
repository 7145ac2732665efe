#!/bin/bash
set -x
O=gpurun_out/r2c2; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
timeout 120 ./tools/micro/dmma_bw > $O/dmma_bw.txt 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-modes > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
tail -5 $O/tests.log; cat $O/dmma_bw.txt
