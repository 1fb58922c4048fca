#!/bin/bash
# perft root chain after dropping the no-op ply launch + programmatic dependent launch: rule tests, probe with / without PDL
mkdir -p gpurun_out
T0=$SECONDS
timeout 600 python -m pytest tests/test_gpu_rules.py tests/test_gpu_fullsize.py -q -x --timeout 300 > gpurun_out/pytest_rules.log 2>&1; echo "== rule tests: $? at $((SECONDS-T0)) s"; tail -3 gpurun_out/pytest_rules.log
echo "--- PDL on"; timeout 120 python scripts/perft_root_probe.py --deep 2>&1 | grep -E " bulk " | grep -E "65536|1048576 " 
echo "--- PDL off"; CRL_NO_PDL=1 timeout 120 python scripts/perft_root_probe.py --deep 2>&1 | grep -E " bulk " | grep -E "65536|1048576 "
echo "--- PDL on again"; timeout 120 python scripts/perft_root_probe.py 2>&1 | grep -E " bulk "
timeout 300 python bench.py --steps 1 --warmup 3 --games 512 --sims 50 --no-cpu-baseline --no-kernels --no-whole-games --no-large > gpurun_out/bench_perft.log 2> gpurun_out/bench_perft.err; echo "== bench perft section: $? at $((SECONDS-T0)) s"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_perft.log").read().strip().splitlines()[-1])
p = d["perft"]
print({k: (p[k]["ms"], round(p[k]["nodes_per_s"] / 1e9, 1)) for k in ("start", "kiwipete")}, round(p["nodes_per_s"] / 1e9, 1), round(p["deep_nodes_per_s"] / 1e9, 1))
PY
