#!/bin/bash
# Round 2, final build: the driver's bench command on 8 GPUs (short legs)
mkdir -p gpurun_out
T0=$SECONDS
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus 8 --steps 3 --warmup 3 --no-kernels --no-training > gpurun_out/bench_8gpu.log 2> gpurun_out/bench_8gpu.err; echo "== 8-GPU bench: $? at $((SECONDS-T0)) s"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_8gpu.log").read().strip().splitlines()[-1])
    print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "n_gpus", d["n_gpus"], d["clocks"])
    print("large", d["large_config"]); print("perft_sharded", {k: (v["ms_max_over_ranks"], round(v["nodes_per_s"] / 1e9, 1)) for k, v in d["perft_sharded"].items()})
except Exception as ex:
    print("ERR", ex); print(open("gpurun_out/bench_8gpu.err").read()[-3000:])
PY
