#!/bin/bash
# Round 2, final build: the new GPU tests, the driver's bench command on 2 GPUs, and a 2-rank selfplay run (games sharded by
# rank, finished games gathered over NCCL) with evaluation reuse on
mkdir -p gpurun_out
T0=$SECONDS
timeout 600 python -m pytest tests/test_gpu_tree.py tests/test_gpu_reuse.py -q -x --timeout 300 > gpurun_out/pytest_tree2.log 2>&1; echo "== tree / reuse tests: $? at $((SECONDS-T0)) s"; tail -3 gpurun_out/pytest_tree2.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu.log 2> gpurun_out/bench_2gpu.err; echo "== 2-GPU bench: $? at $((SECONDS-T0)) s"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_2gpu.log").read().strip().splitlines()[-1])
    print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "n_gpus", d["n_gpus"], d["clocks"])
    print("whole", d["whole_games"]["simulations_per_s"], "reuse", d["whole_games_reuse"]["simulations_per_s"]); print("large", d["large_config"]); print("perft_sharded", d["perft_sharded"])
except Exception as ex:
    print("ERR", ex); print(open("gpurun_out/bench_2gpu.err").read()[-3000:])
PY
rm -rf /tmp/m2; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
  -m chessrl_b200.selfplay /tmp/m2 --games 96 --lanes 24 --sims 40 --max-moves 30 --no-train > gpurun_out/selfplay_2gpu.log 2>&1; echo "== 2-rank selfplay: $? at $((SECONDS-T0)) s"; tail -3 gpurun_out/selfplay_2gpu.log
python -c "
import json; d=json.load(open('/tmp/m2/gameplays.json')); print('games gathered:', len(d), 'plies', sum(len(g['moves']) for g in d))"
