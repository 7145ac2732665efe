#!/bin/bash
O=gpurun_out/r2c37; mkdir -p $O
timeout 1200 python -m pytest tests/test_dropin_gpu.py tests/test_golden_round_gpu.py tests/test_comm_gpu.py -m gpu -q -x > $O/tests.log 2>&1; tail -5 $O/tests.log
for ep in 40 5 1024; do
python tools/rollout_ncu.py 4 $ep > $O/rollout${ep}.log 2>&1; echo "fused env step, $ep episodes:"; tail -1 $O/rollout${ep}.log
MBO_ENV_TORCH=1 python tools/rollout_ncu.py 4 $ep > $O/rollout${ep}_torch.log 2>&1; echo "torch env step, $ep episodes:"; tail -1 $O/rollout${ep}_torch.log
done
