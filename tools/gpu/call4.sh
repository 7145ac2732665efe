#!/bin/bash
# 2 GPUs: data-parallel PPO update, peer-memory exchange vs eager NCCL, and the bench at N=2
set -x
O=gpurun_out/r2c4; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 python tools/ppo_multi_gpu.py > $O/ppo_1gpu.json 2> $O/ppo_1gpu.err
timeout 300 $TR --master-port 29541 tools/ppo_multi_gpu.py --exchange peer > $O/ppo_2gpu_peer.json 2> $O/ppo_2gpu_peer.err
timeout 300 $TR --master-port 29542 tools/ppo_multi_gpu.py --exchange nccl > $O/ppo_2gpu_nccl.json 2> $O/ppo_2gpu_nccl.err
timeout 300 $TR --master-port 29543 tools/ppo_multi_gpu.py --exchange peer --episodes 2048 --epochs 1 > $O/ppo_2gpu_peer_big.json 2> $O/ppo_2gpu_peer_big.err
timeout 600 $TR --master-port 29544 bench.py --gpus 2 --steps 10 --warmup 3 --no-modes > $O/bench_2gpu.json 2> $O/bench_2gpu.err
tail -2 $O/*.json | cut -c1-1500; tail -5 $O/*.err
