#!/bin/bash
# Round 2: GPU suite (sharded root perft, warp-parallel repetition walk) + the 2-GPU bench with the device-driven sharded perft
mkdir -p gpurun_out
T0=$SECONDS
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/pytest_gpu.log 2>&1; echo "== pytest -m gpu: $? at $((SECONDS-T0)) s"; tail -8 gpurun_out/pytest_gpu.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 3 --warmup 3 --no-whole-games --no-large > gpurun_out/bench_2gpu_b.log 2> gpurun_out/bench_2gpu_b.err; echo "== 2-GPU bench: $? at $((SECONDS-T0)) s"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_2gpu_b.log").read().strip().splitlines()[-1])
    print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "n_gpus", d["n_gpus"], d["clocks"])
    for k, v in d["perft_sharded"].items():
        print(k, round(v["ms_max_over_ranks"], 3), "ms", round(v["nodes_per_s"] / 1e9, 1), "G nodes/s", v["lanes_this_rank"], v["breadth_first_plies"])
except Exception as ex:
    print("ERR", ex); print(open("gpurun_out/bench_2gpu_b.err").read()[-3000:])
PY
timeout 600 python bench.py --whole-games 2048 --wg-lanes 512 --wg-sims 50 --steps 2 --warmup 3 --no-cpu-baseline --no-perft --no-kernels --no-whole-games --no-large > gpurun_out/bench_complete_run.log 2>&1; echo "== complete run: $? at $((SECONDS-T0)) s"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_complete_run.log").read().strip().splitlines()[-1])
    print("complete", d["whole_games_complete_run"])
except Exception as ex:
    print("ERR", ex); print(open("gpurun_out/bench_complete_run.log").read()[-3000:])
PY
