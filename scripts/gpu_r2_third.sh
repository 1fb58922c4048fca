#!/bin/bash
# Round 2, third GPU call: warp-cooperative generator in the tree kernels and the small perft plies
mkdir -p gpurun_out
T0=$SECONDS
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/pytest_gpu.log 2>&1; echo "== pytest -m gpu: $? at $((SECONDS-T0)) s"; tail -8 gpurun_out/pytest_gpu.log
S="--no-cpu-baseline --no-perft --no-kernels --no-whole-games --no-large"
timeout 300 python bench.py --steps 4 --warmup 3 $S > gpurun_out/bench_warpgen.log 2>&1; echo "== 4096 lanes: $?"
timeout 300 python bench.py --games 512 --sims 200 --steps 4 --warmup 3 $S > gpurun_out/bench_warpgen_512.log 2>&1; echo "== 512 lanes: $?"
timeout 300 python bench.py --games 512 --sims 200 --inflight 8 --steps 4 --warmup 3 $S > gpurun_out/bench_warpgen_512k8.log 2>&1; echo "== 512 lanes K=8: $?"
python - <<'PY'
import json
for f in ("warpgen", "warpgen_512", "warpgen_512k8"):
    try:
        d = json.loads(open("gpurun_out/bench_%s.log" % f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "roof", round(d["roofline"]["achieved"]), d["roofline"]["share_of_step_ms"], d["clocks"])
    except Exception as ex:
        print(f, "ERR", ex); print(open("gpurun_out/bench_%s.log" % f).read()[-2000:])
PY
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-whole-games --no-large > gpurun_out/bench_perft.log 2>&1; echo "== perft + kernels: $? at $((SECONDS-T0)) s"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_perft.log").read().strip().splitlines()[-1])
    p = d["perft"]
    print("perft start", p["start"]["ms"], round(p["start"]["nodes_per_s"] / 1e9, 1), "kiwi", p["kiwipete"]["ms"], round(p["kiwipete"]["nodes_per_s"] / 1e9, 1), "both", round(p["nodes_per_s"] / 1e9, 1), "deep", round(p["deep_nodes_per_s"] / 1e9, 1))
    print("no-bulk", p["start"]["no_bulk_ms"], p["kiwipete"]["no_bulk_ms"], "deep", p["start_d7"], p["kiwipete_d6"])
except Exception as ex:
    print("ERR", ex); print(open("gpurun_out/bench_perft.log").read()[-2000:])
PY
B="--games 4096 --sims 6 --steps 1 --warmup 1 $S"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' -c 900 --csv --log-file gpurun_out/launches_step_r02c.csv \
   python bench.py $B > gpurun_out/ncu_launches_r02c.log 2>&1; echo "== launch list: $? at $((SECONDS-T0)) s"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' -c 400 --csv --log-file gpurun_out/launches_perft_r02c.csv \
   python scripts/perft_root_probe.py > gpurun_out/ncu_launches_perft_r02c.log 2>&1; echo "== perft launch list: $? at $((SECONDS-T0)) s"
export CRL_NO_GRAPH=1
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_$tool.log 2>&1
  echo "== smoke $tool: $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok" gpurun_out/sanitize_$tool.log | tail -3
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/pair_sanitize.py > gpurun_out/sanitize_perft_$tool.log 2>&1
  echo "== perft $tool: $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok" gpurun_out/sanitize_perft_$tool.log | tail -4
done
ls -la gpurun_out | tail -8
