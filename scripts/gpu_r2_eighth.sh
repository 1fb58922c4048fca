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
