#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_api.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_api.log 2>&1; echo "== api: exit $?"; tail -25 gpurun_out/pytest_api.log
# launch list of one lockstep step (my kernels only; skips the synthetic-position setup)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_(select|reply|final|encode_rows|conv|heads|root)' -s 100 -c 500 --csv --log-file gpurun_out/launches_step.csv \
   python bench.py --games 4096 --sims 6 --steps 1 --warmup 1 --no-cpu-baseline --no-perft > gpurun_out/ncu_launches2.log 2>&1; echo "== ncu launches: exit $?"
