#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rules.py tests/test_gpu_tree.py tests/test_gpu_fullsize.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_rules.log 2>&1; echo "tests $?"; tail -3 gpurun_out/pytest_rules.log
timeout 600 python scripts/perft_probe.py --time > gpurun_out/perft_probe_new.log 2>&1; echo "probe $?"; cat gpurun_out/perft_probe_new.log
CRL_PERFT_6BLOCKS=1 timeout 600 python scripts/perft_probe.py --time > gpurun_out/perft_probe_6b.log 2>&1; echo "probe 6 blocks $?"; grep -E "perft" gpurun_out/perft_probe_6b.log
