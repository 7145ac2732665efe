#!/bin/bash
# N GPUs (second pass of round 2): bench incl. ppo_train, data-parallel PPO at the reference batch and at 1024 episodes / GPU, sweep
N=$1
set -x
O=gpurun_out/r2fin3_n$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29551 bench.py --gpus $N --steps 20 --warmup 5 --no-modes > $O/bench.json 2> $O/bench.err
timeout 300 $TR --master-port 29552 tools/ppo_multi_gpu.py --exchange peer > $O/ppo_peer.json 2> $O/ppo_peer.err
timeout 300 $TR --master-port 29555 tools/ppo_multi_gpu.py --exchange peer --episodes $((1024*N)) --epochs 1 > $O/ppo_peer_big.json 2> $O/ppo_peer_big.err
timeout 600 $TR --master-port 29556 tools/sweep.py > $O/sweep.md 2> $O/sweep.err
for f in $O/*.json; do echo "== $f"; tail -n 1 $f | cut -c1-1200; done
tail -n 14 $O/sweep.md
grep -v "OMP_NUM\|^\*\*\*" $O/*.err | tail -n 20
