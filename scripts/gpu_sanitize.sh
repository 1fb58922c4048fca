#!/bin/bash
# compute-sanitizer over the smoke test (rules, tree search with both evaluators, the tcgen05 tower)
mkdir -p gpurun_out
export CRL_NO_GRAPH=1
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool: $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok" gpurun_out/sanitize_$tool.log | tail -3
done
