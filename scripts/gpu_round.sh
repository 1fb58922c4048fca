#!/bin/bash
# one GPU: all parity tests, smoke, the default bench line, then the large configurations
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "== pytest -m gpu: $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke: $?"; tail -2 gpurun_out/smoke.log
show() { python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$1.log").read().strip().splitlines()[-1])
    print("$1", round(d["value"]), "e2e", round(d["e2e"]["value"]), "roof", round(d["roofline"]["achieved"]), round(d["roofline"]["frac"],3), d["clocks"])
    for k,v in (d.get("kernels") or {}).items(): print("   ", k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a not in ("workload","bound")})
    if d.get("perft"): print("    perft", {k:(v if not isinstance(v,dict) else {a:b for a,b in v.items() if a in ("nodes_per_s","dfs_only_no_bulk_nodes_per_s")}) for k,v in d["perft"].items()})
    print("    cpu", d.get("cpu_baseline"))
except Exception as ex:
    print("$1 ERR", ex); print(open("gpurun_out/$1.err").read()[-2000:])
PY
}
timeout 900 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "== default bench: $?"; show bench_default
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2> gpurun_out/bench_reference.err; echo "== reference arm: $?"; tail -c 600 gpurun_out/bench_reference.log
if [ "$BIG" = "1" ]; then
timeout 900 python bench.py --games 8192 --sims 800 --steps 1 --warmup 1 --no-cpu-baseline --no-perft --no-kernels > gpurun_out/bench_8192x800.log 2> gpurun_out/bench_8192x800.err; echo "== 8192x800: $?"; show bench_8192x800
timeout 1200 python bench.py --games 65536 --sims 200 --steps 1 --warmup 1 --no-cpu-baseline --no-perft --no-kernels > gpurun_out/bench_65536x200.log 2> gpurun_out/bench_65536x200.err; echo "== 65536x200: $?"; show bench_65536x200
fi
