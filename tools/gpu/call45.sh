#!/bin/bash
O=gpurun_out/r2c45; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/tests.log 2>&1; tail -3 $O/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 300 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; tail -c 300 $O/bench.json
