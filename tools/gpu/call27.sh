#!/bin/bash
O=gpurun_out/r2c27; mkdir -p $O
for f in 0 1; do
MBO_FUSE_TOPK=$f python bench.py --steps 30 --warmup 5 --no-modes --no-cpu-baseline --no-ppo > $O/bench_topk$f.json 2> $O/bench_topk$f.err
python -c "
import json,sys
d=json.loads(open('$O/bench_topk$f.json').read().strip().splitlines()[-1])
print('fuse_topk=$f ms_per_step', d['ms_per_step'], 'launches', d['gpu_launches'], 'mlp ms', d['roofline']['avg_launch_ms'])
"
done
