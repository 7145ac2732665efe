#!/bin/bash
O=gpurun_out/r2c12; mkdir -p $O
timeout 200 python tools/overlap_probe.py fp16 > $O/overlap.log 2>&1; cat $O/overlap.log
timeout 300 python tools/tc_probe.py > $O/tc_probe.log 2>&1; tail -8 $O/tc_probe.log
timeout 600 python -m pytest tests/test_af_gpu.py -m gpu -q -x > $O/tests.log 2>&1; tail -3 $O/tests.log
