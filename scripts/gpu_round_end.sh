#!/bin/bash
# Round-end validation on one B200, the driver's own sequence: GPU suite, smoke, the default bench command line, the reference arm
mkdir -p gpurun_out
T0=$SECONDS
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "== pytest -m gpu: $? at $((SECONDS-T0)) s"; tail -4 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke: $? at $((SECONDS-T0)) s"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_round_end.log 2> gpurun_out/bench_round_end.err; echo "== bench --steps 20 --warmup 5: $? at $((SECONDS-T0)) s"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_round_end.log").read().strip().splitlines()[-1])
    print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "roof", round(d["roofline"]["achieved"]), round(d["roofline"]["frac"], 3), "traffic", d["roofline"]["traffic"], d["clocks"])
    print("perft", round(d["perft"]["nodes_per_s"] / 1e9, 1), d["perft"]["start"]["ms"], d["perft"]["kiwipete"]["ms"], "deep", round(d["perft"]["deep_nodes_per_s"] / 1e9, 1), "cpu", round(d["cpu_baseline"]["value"], 1), d["cpu_baseline"]["cores"])
    print("whole", round(d["whole_games"]["simulations_per_s"]), d["whole_games"]["fraction_of_device_resident_value"], "large", round(d["large_config"]["simulations_per_s"]))
    print("launches", d["gpu_launches"], "ms/step", d["ms_per_step"])
except Exception as ex:
    print("ERR", ex); print(open("gpurun_out/bench_round_end.err").read()[-3000:])
PY
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_reference_arm.log 2>&1; echo "== reference arm: $? at $((SECONDS-T0)) s"; tail -1 gpurun_out/bench_reference_arm.log | cut -c1-400
