#!/bin/bash
mkdir -p gpurun_out/r2c11
timeout 120 ./tools/micro/coresidency > gpurun_out/r2c11/coresidency.txt 2>&1
tail -3 gpurun_out/r2c11/coresidency.txt
PROBE_PRIO=0 timeout 200 python tools/overlap_probe.py fp16
