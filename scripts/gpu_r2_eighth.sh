#!/bin/bash
# all GPU tests on the current build, smoke, then compute-sanitizer (memcheck, racecheck) over a small self-play run
# with evaluation reuse / the fused head / the statistics-only softmax
mkdir -p gpurun_out
T0=$SECONDS
timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "== pytest -m gpu: $? at $((SECONDS-T0)) s"; tail -4 gpurun_out/pytest_gpu.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke: $? at $((SECONDS-T0)) s"; tail -1 gpurun_out/smoke.log
export CRL_NO_GRAPH=1
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/reuse_sanitize.py > gpurun_out/sanitize_reuse_$tool.log 2>&1
  echo "== $tool: $? at $((SECONDS-T0)) s"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|evaluator:|network:|workload ok" gpurun_out/sanitize_reuse_$tool.log | tail -5
done
unset CRL_NO_GRAPH
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_(bfs_first|bfs_ply|perft_walk)' -c 60 --csv --log-file gpurun_out/launches_perft.csv \
   python scripts/perft_root_probe.py > gpurun_out/ncu_perft.log 2>&1; echo "== ncu perft launch list: $? at $((SECONDS-T0)) s"
python - <<'PY'
import csv
rows = [r for r in csv.reader(l for l in open("gpurun_out/launches_perft.csv") if not l.startswith("=="))]
h = rows[0]; kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
for r in rows[1:21]:
    print(r[kn].split("(")[0][:30], r[mv], r[mu])
PY
