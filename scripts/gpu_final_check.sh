#!/bin/bash
# last check of a round: whole GPU suite, smoke, the default bench line
mkdir -p gpurun_out
T0=$SECONDS
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "== pytest -m gpu: $? at $((SECONDS-T0)) s"; tail -3 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke: $? at $((SECONDS-T0)) s"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "== default bench: $? at $((SECONDS-T0)) s"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_default.log").read().strip().splitlines()[-1])
    print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "roof", round(d["roofline"]["achieved"]), round(d["roofline"]["frac"], 3), d["clocks"], d["roofline"]["share_of_step_ms"])
    print("perft", d["perft"]["start"]["ms"], d["perft"]["kiwipete"]["ms"], round(d["perft"]["nodes_per_s"] / 1e9, 1), "deep", round(d["perft"]["deep_nodes_per_s"] / 1e9, 1), "cpu", round(d["cpu_baseline"]["value"], 1))
    print("whole", round(d["whole_games"]["simulations_per_s"]), "reuse", round(d["whole_games_reuse"]["simulations_per_s"]), "large", round(d["large_config"]["simulations_per_s"]), "train", {k: round(v["positions_per_s"]) for k, v in d["training_step"]["by_precision"].items()})
except Exception as ex:
    print("ERR", ex); print(open("gpurun_out/bench_default.err").read()[-2000:])
PY
