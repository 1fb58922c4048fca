#!/bin/bash
# N GPUs of one box (gpurun --gpus N): the default bench line (4,096 games x 200 sims per GPU) and BASELINE configs[4]
# (65,536 games x 800 sims/move sharded by game: 65,536/N games per GPU), launched exactly as the driver does.
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571"
timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/scale${N}_default_v4.log 2> gpurun_out/scale${N}_default_v4.err; echo "== default x$N: $?"
tail -1 gpurun_out/scale${N}_default_v4.log | cut -c1-400
G=$((65536 / N))
timeout 1500 $TR bench.py --gpus $N --games $G --sims 800 --steps 1 --warmup 3 --no-kernels > gpurun_out/scale${N}_65536x800_v4.log 2> gpurun_out/scale${N}_65536x800_v4.err; echo "== 65536x800 over $N: $?"
tail -1 gpurun_out/scale${N}_65536x800_v4.log | cut -c1-400
nvidia-smi --query-gpu=index,name,clocks.sm,power.draw --format=csv,noheader | head -8
