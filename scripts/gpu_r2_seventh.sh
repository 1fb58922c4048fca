#!/bin/bash
# reuse tests after the warp-race fix + an ncu launch list of the step kernels (CRL_NO_GRAPH so that each launch is visible)
mkdir -p gpurun_out
T0=$SECONDS
timeout 600 python -m pytest tests/test_gpu_reuse.py tests/test_gpu_tree.py tests/test_gpu_parity_net.py -q --timeout 300 > gpurun_out/pytest_reuse.log 2>&1; echo "== reuse/tree tests: $? at $((SECONDS-T0)) s"; tail -6 gpurun_out/pytest_reuse.log
S="--no-cpu-baseline --no-perft --no-kernels --no-whole-games --no-large"
CRL_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_(select_expand|reply|finalize|encode_rows|softmax_value|conv_v2|trunk4)' -s 60 -c 240 --csv --log-file gpurun_out/launches_seventh.csv \
   python bench.py --games 4096 --sims 12 --steps 1 --warmup 1 $S > gpurun_out/ncu_seventh.log 2>&1; echo "== ncu launch list: $? at $((SECONDS-T0)) s"
python scripts/ncu_summarise.py launches gpurun_out/launches_seventh.csv | tail -12
