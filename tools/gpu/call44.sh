#!/bin/bash
O=gpurun_out/r2c44; mkdir -p $O
timeout 600 python -m pytest tests/test_af_gpu.py tests/test_golden_round_gpu.py tests/test_ppo_update_simt_gpu.py -m gpu -q -x > $O/tests.log 2>&1; tail -3 $O/tests.log
timeout 300 python tools/run_configs.py C4 > $O/c4.jsonl 2> $O/c4.err
python - <<'PY'
import json
for l in open('gpurun_out/r2c44/c4.jsonl'):
    c=json.loads(l); print(c['config'][:70], {k:(round(v,4) if isinstance(v,float) else v) for k,v in c.items() if k in ('t_batch_s','env_steps_per_s','steady_state_e2e_steps_per_s','rollout_sps')})
PY
