#!/bin/bash
mkdir -p gpurun_out
for f in rules encode tree net api parity_net; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -m gpu -q -x --timeout 800 > gpurun_out/pytest_$f.log 2>&1
  echo "== $f: exit $?"; tail -4 gpurun_out/pytest_$f.log
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke: exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py ${BENCH_ARGS} > gpurun_out/bench_latest.log 2> gpurun_out/bench_latest.err; echo "== bench: exit $?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/bench_latest.log").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["share_of_step_ms"], d["clocks"], "cpu", d.get("cpu_baseline"), "perft", d.get("perft"))
except Exception as ex:
    print("ERR", ex); print(open("gpurun_out/bench_latest.log").read()[-2000:]); print(open("gpurun_out/bench_latest.err").read()[-2000:])
PY
