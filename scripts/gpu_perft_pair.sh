#!/bin/bash
mkdir -p gpurun_out
T0=$SECONDS
timeout 120 python -m pytest tests/test_gpu_rules.py -m gpu -q -x --timeout 100 > gpurun_out/pytest_rules_pair6.log 2>&1; echo "== rules (pair=6): $? at $((SECONDS-T0)) s"; tail -4 gpurun_out/pytest_rules_pair6.log
for m in 5 0; do
CRL_PERFT_PAIR=$m timeout 90 python -m pytest tests/test_gpu_rules.py -m gpu -q -x --timeout 80 -k "perft_root or corner" > gpurun_out/pytest_rules_pair$m.log 2>&1; echo "== perft_root tests (pair=$m): $? at $((SECONDS-T0)) s"; tail -3 gpurun_out/pytest_rules_pair$m.log
done
timeout 120 python scripts/perft_pair_probe.py > gpurun_out/perft_pair_probe.log 2>&1; echo "== probe: $? at $((SECONDS-T0)) s"; tail -32 gpurun_out/perft_pair_probe.log
