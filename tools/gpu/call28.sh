#!/bin/bash
O=gpurun_out/r2c28; mkdir -p $O
for m in tf32 bf16x3; do
for f in 0 1; do
MBO_FUSE_TOPK=$f python bench.py --steps 15 --warmup 4 --mlp $m --no-modes --no-cpu-baseline --no-ppo > $O/bench_${m}_topk$f.json 2> $O/bench_${m}_topk$f.err
python -c "
import json,sys
d=json.loads(open('$O/bench_${m}_topk$f.json').read().strip().splitlines()[-1])
print('$m fuse_topk=$f ms_per_step', d['ms_per_step'], 'mlp ms', d['roofline']['avg_launch_ms'])
"
done
done
