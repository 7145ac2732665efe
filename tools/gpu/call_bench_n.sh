#!/bin/bash
# bench.py only, on N GPUs (final build)
N=$1
O=gpurun_out/r2fin3_n$N; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 20 --warmup 5 --no-modes > $O/bench.json 2> $O/bench.err
tail -n 1 $O/bench.json | cut -c1-300
