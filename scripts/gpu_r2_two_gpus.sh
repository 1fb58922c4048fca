#!/bin/bash
# Round 2: the driver's multi-GPU bench command on 2 GPUs (weak scaling legs + sharded perft + BASELINE configs[4] leg)
mkdir -p gpurun_out
T0=$SECONDS
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu.log 2> gpurun_out/bench_2gpu.err; echo "== 2-GPU bench: $? at $((SECONDS-T0)) s"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_2gpu.log").read().strip().splitlines()[-1])
    print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "n_gpus", d["n_gpus"], d["clocks"])
    print("whole", d["whole_games"]); print("large", d["large_config"]); print("perft_sharded", d["perft_sharded"])
except Exception as ex:
    print("ERR", ex); print(open("gpurun_out/bench_2gpu.err").read()[-3000:])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_2gpu_ref.log 2>&1; echo "== reference arm under torchrun: $? at $((SECONDS-T0)) s"; tail -1 gpurun_out/bench_2gpu_ref.log | cut -c1-300
