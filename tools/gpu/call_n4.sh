#!/bin/bash
set -x
O=gpurun_out/r2n4; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
( time timeout 600 $TR --master-port 29561 bench.py --gpus 4 --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err ) 2> $O/bench_time.txt
( time timeout 600 $TR --master-port 29562 bench.py --impl reference --gpus 4 --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err ) 2> $O/bench_ref_time.txt
timeout 300 $TR --master-port 29563 tools/ppo_multi_gpu.py --exchange peer > $O/ppo_peer.json 2> $O/ppo_peer.err
tail -n 1 $O/bench.json | cut -c1-300; tail -n 1 $O/bench_ref.json | cut -c1-200; tail -n 1 $O/ppo_peer.json | cut -c1-500; cat $O/bench_time.txt $O/bench_ref_time.txt | grep real
grep -v "OMP_NUM\|^\*\*\*" $O/*.err | tail -n 10
