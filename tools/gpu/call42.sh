#!/bin/bash
O=gpurun_out/r2c42; mkdir -p $O
timeout 600 python -m pytest tests/test_gp_gpu.py tests/test_golden_round_gpu.py -m gpu -q -x > $O/tests.log 2>&1; tail -3 $O/tests.log
echo "register k* for n <= 64:"; python tools/gp_probe.py 2>&1 | grep "n= 48\|n= 64\|n= 30\|n=100"
echo "shared-memory k* for n > 32 (MBO_GP_NO_WIDE_REG=1):"; MBO_GP_NO_WIDE_REG=1 python tools/gp_probe.py 2>&1 | grep "n= 48\|n= 64"
