#!/bin/bash
O=gpurun_out/r2c16; mkdir -p $O
timeout 300 python tools/gp_probe.py > $O/gp_probe.log 2>&1; cat $O/gp_probe.log
timeout 900 python -m pytest tests/test_gp_gpu.py tests/test_comm_gpu.py tests/test_golden_round_gpu.py -m gpu -q > $O/tests.log 2>&1; tail -3 $O/tests.log
