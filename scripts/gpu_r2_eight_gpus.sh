#!/bin/bash
# Round 2: the driver's multi-GPU bench command on 8 GPUs (weak-scaling legs, sharded perft, BASELINE configs[4] leg)
mkdir -p gpurun_out
T0=$SECONDS
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_8gpu.log 2> gpurun_out/bench_8gpu.err; echo "== 8-GPU bench: $? at $((SECONDS-T0)) s"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_8gpu.log").read().strip().splitlines()[-1])
    print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "n_gpus", d["n_gpus"], d["clocks"])
    print("whole", d["whole_games"]["simulations_per_s"], "large", d["large_config"])
    for k, v in d["perft_sharded"].items():
        print(k, round(v["ms_max_over_ranks"], 3), "ms", round(v["nodes_per_s"] / 1e9, 1), "G nodes/s", v["lanes_this_rank"], v["breadth_first_plies"])
except Exception as ex:
    print("ERR", ex); print(open("gpurun_out/bench_8gpu.err").read()[-3000:])
PY
