#!/bin/bash
# the default bench line of the current build (what the driver runs), then a complete finite run with and without reuse
mkdir -p gpurun_out
T0=$SECONDS
timeout 600 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "== default bench: $? at $((SECONDS-T0)) s"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_default.log").read().strip().splitlines()[-1])
    print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "roof", round(d["roofline"]["achieved"]), round(d["roofline"]["frac"], 3), d["clocks"], d["roofline"]["share_of_step_ms"])
    print("perft", d["perft"]["start"]["ms"], d["perft"]["kiwipete"]["ms"], round(d["perft"]["nodes_per_s"] / 1e9, 1), "deep", round(d["perft"]["deep_nodes_per_s"] / 1e9, 1), "cpu", round(d["cpu_baseline"]["value"], 1), d["cpu_baseline"]["cores"])
    for k in ("whole_games", "whole_games_reuse", "large_config"):
        print(k, {a: b for a, b in d[k].items() if a not in ("workload", "timing")})
except Exception as ex:
    print("ERR", ex); print(open("gpurun_out/bench_default.err").read()[-1500:])
PY
for r in "" "--no-reuse"; do
timeout 300 python -m chessrl_b200.selfplay /tmp/models$r --games 2048 --lanes 512 --sims 50 --no-train $r > gpurun_out/selfplay_complete$r.log 2>&1; echo "== complete run $r: $? at $((SECONDS-T0)) s"; tail -2 gpurun_out/selfplay_complete$r.log
done
