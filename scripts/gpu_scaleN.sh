#!/bin/bash
# N GPUs of one box: the default bench line (4,096 games x 200 sims per GPU), launched exactly as the driver does
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571"
timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/scale${N}_default_v4.log 2> gpurun_out/scale${N}_default_v4.err; echo "== default x$N: $?"
tail -1 gpurun_out/scale${N}_default_v4.log | cut -c1-300
