#!/bin/bash
# compute-sanitizer over crl_perft_root_host (default path and the optional two-ply pass k_perft_pair)
mkdir -p gpurun_out
T0=$SECONDS
for tool in memcheck racecheck; do
  timeout 55 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/pair_sanitize.py > gpurun_out/pair_sanitize_$tool.log 2>&1
  echo "== $tool: $? at $((SECONDS-T0)) s"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|pair mode" gpurun_out/pair_sanitize_$tool.log | tail -5
done
