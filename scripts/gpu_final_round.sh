#!/bin/bash
# Round-end validation on one B200: all GPU parity tests, smoke, the default bench line, the reference arm is CPU-only
# (run in the build container), then two sizing probes (lanes = a multiple of one persistent-grid round of 592).
mkdir -p gpurun_out
T0=$SECONDS
timeout 200 python -m pytest tests -m gpu -q -x --timeout 150 > gpurun_out/pytest_gpu.log 2>&1; echo "== pytest -m gpu: $? at $((SECONDS-T0)) s"; tail -3 gpurun_out/pytest_gpu.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke: $? at $((SECONDS-T0)) s"; tail -1 gpurun_out/smoke.log
timeout 240 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "== default bench: $? at $((SECONDS-T0)) s"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_default.log").read().strip().splitlines()[-1])
    print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "roof", round(d["roofline"]["achieved"]), round(d["roofline"]["frac"], 3), d["clocks"])
    print("perft", round(d["perft"]["nodes_per_s"] / 1e9, 1), "deep", round(d["perft"]["deep_nodes_per_s"] / 1e9, 1), "cpu", round(d["cpu_baseline"]["value"], 1), d["cpu_baseline"]["cores"])
except Exception as ex:
    print("ERR", ex); print(open("gpurun_out/bench_default.err").read()[-1500:])
PY
for g in 4144 4096; do
timeout 90 python bench.py --games $g --sims 200 --steps 2 --warmup 3 --no-cpu-baseline --no-perft --no-kernels > gpurun_out/bench_$g.log 2> gpurun_out/bench_$g.err; echo "== $g lanes: $? at $((SECONDS-T0)) s"
python -c "
import json
d=json.loads(open('gpurun_out/bench_$g.log').read().strip().splitlines()[-1]); print($g, round(d['value']), 'e2e', round(d['e2e']['value']), d['clocks'])"
done
