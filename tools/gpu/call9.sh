#!/bin/bash
set -x
O=gpurun_out/r2c10; mkdir -p $O
timeout 200 python tools/overlap_probe.py fp16 >> $O/overlap.log 2>&1
cat $O/overlap.log
