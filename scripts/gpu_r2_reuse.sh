#!/bin/bash
# Evaluation reuse on the device: the A/B parity tests, the tree / API suites that now run with reuse on by default,
# and the steady-state whole-game leg with and without reuse.
mkdir -p gpurun_out
T0=$SECONDS
timeout 300 python -m pytest tests/test_gpu_reuse.py -q -x --timeout 200 > gpurun_out/pytest_reuse.log 2>&1; echo "== reuse tests: $? at $((SECONDS-T0)) s"; tail -15 gpurun_out/pytest_reuse.log
timeout 400 python -m pytest tests/test_gpu_tree.py tests/test_gpu_api.py tests/test_gpu_parity_net.py -q -x --timeout 200 > gpurun_out/pytest_tree_api.log 2>&1; echo "== tree/api tests: $? at $((SECONDS-T0)) s"; tail -5 gpurun_out/pytest_tree_api.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-perft --no-kernels --no-large > gpurun_out/bench_reuse.log 2> gpurun_out/bench_reuse.err; echo "== bench: $? at $((SECONDS-T0)) s"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_reuse.log").read().strip().splitlines()[-1])
    print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["clocks"])
    print("whole", json.dumps(d["whole_games"]))
    print("whole_reuse", json.dumps(d["whole_games_reuse"]))
except Exception as ex:
    print("ERR", ex); print(open("gpurun_out/bench_reuse.err").read()[-2500:])
PY
