#!/bin/bash
mkdir -p gpurun_out
for f in net tree api; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_v2_$f.log 2>&1
  echo "== $f: exit $?"; tail -4 gpurun_out/pytest_v2_$f.log
done
B="--games 4096 --sims 200 --steps 2 --warmup 3 --no-cpu-baseline --no-perft"
timeout 600 python bench.py $B > gpurun_out/bench_v2.log 2>&1; echo "== v2+graph: $?"; python - <<'PY'
import json
for n in ["bench_v2"]:
    try:
        d=json.loads(open("gpurun_out/%s.log"%n).read().strip().splitlines()[-1]); print(n, d["value"], d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["share_of_step_ms"], d["clocks"])
    except Exception as ex: print(n, "ERR", ex); print(open("gpurun_out/%s.log"%n).read()[-1500:])
PY
CRL_NO_GRAPH=1 timeout 600 python bench.py $B > gpurun_out/bench_v2_nograph.log 2>&1; echo "== v2 nograph: $?"
CRL_CONV_V1=1 timeout 600 python bench.py $B > gpurun_out/bench_v1_graph.log 2>&1; echo "== v1 graph: $?"
python - <<'PY'
import json
for n in ["bench_v2_nograph","bench_v1_graph"]:
    try:
        d=json.loads(open("gpurun_out/%s.log"%n).read().strip().splitlines()[-1]); print(n, d["value"], d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["share_of_step_ms"])
    except Exception as ex: print(n, "ERR", ex); print(open("gpurun_out/%s.log"%n).read()[-1500:])
PY
