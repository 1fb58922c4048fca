#!/bin/bash
# BASELINE configs[0] on the GPU: ONE game, 100 simulations per move (the reference's own CPU-runnable case), exact schedule
# and the reference's default --threads 6 (wave schedule); plus 1 game x 900 (selfplay.py's hard-coded 900)
mkdir -p gpurun_out
S="--no-cpu-baseline --no-perft --no-kernels --no-whole-games --no-large --no-training"
for cfg in "1 100 1" "1 100 6" "1 900 1" "1 900 6" "64 100 1" "296 100 1" "592 100 1"; do
  set -- $cfg
  timeout 200 python bench.py --games $1 --sims $2 --inflight $3 --steps 5 --warmup 3 $S > gpurun_out/b1.log 2> gpurun_out/b1.err
  python - "$cfg" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/b1.log").read().strip().splitlines()[-1])
    print(sys.argv[1], "-> %.0f simulations/s, %.2f ms per move search, e2e %.0f, evals/sim %.2f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["evaluations_per_simulation"]))
except Exception as ex:
    print(sys.argv[1], "ERR", ex, open("gpurun_out/b1.err").read()[-800:])
PY
done
