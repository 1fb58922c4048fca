#!/bin/bash
mkdir -p gpurun_out
for f in net parity_net tree; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -m gpu -q -x --timeout 800 > gpurun_out/pytest_$f.log 2>&1
  echo "== $f: exit $?"; tail -6 gpurun_out/pytest_$f.log
done
B="--games 4096 --sims 200 --steps 2 --warmup 3 --no-cpu-baseline --no-perft"
timeout 600 python bench.py $B > gpurun_out/bench_trunk.log 2>&1; echo "== trunk: $?"
CRL_NO_TRUNK=1 timeout 600 python bench.py $B > gpurun_out/bench_notrunk.log 2>&1; echo "== no trunk: $?"
python - <<'PY'
import json
for n in ["bench_trunk","bench_notrunk"]:
    try:
        d=json.loads(open("gpurun_out/%s.log"%n).read().strip().splitlines()[-1]); print(n, d["value"], d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["share_of_step_ms"], d["clocks"])
    except Exception as ex: print(n, "ERR", ex); print(open("gpurun_out/%s.log"%n).read()[-1500:])
PY
