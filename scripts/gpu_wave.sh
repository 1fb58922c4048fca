#!/bin/bash
# wave mode (K in-flight simulations per game): parity tests, then throughput at small and full lane counts
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_tree.py tests/test_gpu_api.py -m gpu -q -x --timeout 900 > gpurun_out/pytest_wave.log 2>&1; echo "== tree+api: $?"; tail -15 gpurun_out/pytest_wave.log
B="--steps 2 --warmup 3 --no-cpu-baseline --no-perft"
run() { name=$1; shift; timeout 600 python bench.py $B "$@" > gpurun_out/bench_$name.log 2> gpurun_out/bench_$name.err; echo "== $name: $?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$name.log").read().strip().splitlines()[-1])
    print("$name", round(d["value"]), "e2e", round(d["e2e"]["value"]), "evals/sim", round(d["evaluations_per_simulation"],3), "net TF", round(d["net_tflops_in_step"]), d["roofline"]["share_of_step_ms"], d["clocks"]["sm_mhz"])
except Exception as ex:
    print("$name ERR", ex); print(open("gpurun_out/bench_$name.err").read()[-1500:])
PY
}
run g512_k1 --games 512 --sims 200 --inflight 1
run g512_k8 --games 512 --sims 200 --inflight 8
run g256_k16 --games 256 --sims 200 --inflight 16
run g4096_k6 --games 4096 --sims 200 --inflight 6
run g4096_k1 --games 4096 --sims 200 --inflight 1
