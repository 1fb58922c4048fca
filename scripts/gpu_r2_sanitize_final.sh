#!/bin/bash
# compute-sanitizer on the final build: smoke (rules, tree search with both evaluators, the tower incl. its single-tile
# instantiation), the reuse workload, the perft root chain with dependent launches
mkdir -p gpurun_out
export CRL_NO_GRAPH=1
T0=$SECONDS
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_final_smoke_memcheck.log 2>&1
echo "== memcheck smoke: $? at $((SECONDS-T0)) s"; grep -E "ERROR SUMMARY|smoke ok" gpurun_out/sanitize_final_smoke_memcheck.log | tail -2
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/reuse_sanitize.py > gpurun_out/sanitize_final_reuse_memcheck.log 2>&1
echo "== memcheck reuse workload: $? at $((SECONDS-T0)) s"; grep -E "ERROR SUMMARY|workload ok" gpurun_out/sanitize_final_reuse_memcheck.log | tail -2
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/pair_sanitize.py > gpurun_out/sanitize_final_perft_memcheck.log 2>&1
echo "== memcheck perft root: $? at $((SECONDS-T0)) s"; grep -E "ERROR SUMMARY|pair mode" gpurun_out/sanitize_final_perft_memcheck.log | tail -4
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_final_smoke_racecheck.log 2>&1
echo "== racecheck smoke: $? at $((SECONDS-T0)) s"; grep -E "RACECHECK SUMMARY|smoke ok" gpurun_out/sanitize_final_smoke_racecheck.log | tail -2
