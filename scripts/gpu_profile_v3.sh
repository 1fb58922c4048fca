#!/bin/bash
# ncu evidence for the v3 step (1 GPU): launch list of the bench command, --set full on the tower kernel and the
# HBM-bound kernels.  Graph replay is disabled for the --set full passes so kernels appear as plain launches.
mkdir -p gpurun_out
B="--games 4096 --sims 4 --steps 1 --warmup 1 --no-cpu-baseline --no-perft"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v3.csv \
   python bench.py $B > gpurun_out/ncu_launches_v3.log 2>&1; echo "== launch list: $?"
export CRL_NO_GRAPH=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trunk -s 4 -c 2 -o gpurun_out/prof_trunk_v3 \
   python bench.py $B > gpurun_out/ncu_trunk.log 2>&1; echo "== trunk: $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_(select_expand|reply|finalize|encode_rows|softmax_value|conv_v2)' -s 12 -c 12 -o gpurun_out/prof_step_v3 \
   python bench.py $B > gpurun_out/ncu_step.log 2>&1; echo "== step kernels: $?"
ls -la gpurun_out
