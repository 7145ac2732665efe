#!/bin/bash
O=gpurun_out/r2c39; mkdir -p $O
timeout 1200 python -m pytest tests/test_dropin_gpu.py tests/test_af_gpu.py tests/test_golden_round_gpu.py tests/test_comm_gpu.py -m gpu -q -x > $O/tests.log 2>&1; tail -3 $O/tests.log
for ep in 40 5; do
for vb in 64 1; do
MBO_VALUE_TC_MIN_B=$vb python tools/rollout_ncu.py 4 $ep > $O/rollout${ep}_$vb.log 2>&1; echo "state_pack + env_step, $ep episodes, value TC min B $vb:"; tail -1 $O/rollout${ep}_$vb.log
done
done
python tools/rollout_ncu.py 4 1024 > $O/rollout1024.log 2>&1; tail -1 $O/rollout1024.log
