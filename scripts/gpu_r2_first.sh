#!/bin/bash
# Round 2, first GPU call: the whole GPU suite, smoke, the parity calibration numbers, the default bench line, and
# the ncu evidence for the slot-indexed tower (launch list + --set full on k_trunk4).
mkdir -p gpurun_out
T0=$SECONDS
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "== pytest -m gpu: $? at $((SECONDS-T0)) s"; tail -15 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke: $? at $((SECONDS-T0)) s"; tail -1 gpurun_out/smoke.log
timeout 300 python scripts/net_parity_calibrate.py > gpurun_out/net_parity_calibrate.txt 2>&1; echo "== calibrate: $? at $((SECONDS-T0)) s"; cat gpurun_out/net_parity_calibrate.txt
timeout 600 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "== default bench: $? at $((SECONDS-T0)) s"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_default.log").read().strip().splitlines()[-1])
    print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "roof", round(d["roofline"]["achieved"]), round(d["roofline"]["frac"], 3), d["clocks"])
    print("perft", round(d["perft"]["nodes_per_s"] / 1e9, 1), "deep", round(d["perft"]["deep_nodes_per_s"] / 1e9, 1), "cpu", round(d["cpu_baseline"]["value"], 1), d["cpu_baseline"]["cores"])
    print("whole", d["whole_games"]); print("large", d["large_config"])
except Exception as ex:
    print("ERR", ex); print(open("gpurun_out/bench_default.err").read()[-3000:])
PY
B="--games 4096 --sims 6 --steps 1 --warmup 1 --no-cpu-baseline --no-perft --no-kernels --no-whole-games --no-large"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' -c 900 --csv --log-file gpurun_out/launches_step_r02.csv \
   python bench.py $B > gpurun_out/ncu_launches_r02.log 2>&1; echo "== launch list: $? at $((SECONDS-T0)) s"
CRL_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trunk4 -s 4 -c 2 -o gpurun_out/prof_trunk_r02 \
   python bench.py $B > gpurun_out/ncu_trunk_r02.log 2>&1; echo "== trunk4 full: $? at $((SECONDS-T0)) s"
ls -la gpurun_out | tail -12
