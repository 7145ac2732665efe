#!/bin/bash
set -x
O=gpurun_out/r2c5; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
grep -n "^FAILED\|^ERROR\|passed\|failed" $O/tests.log | tail -20
