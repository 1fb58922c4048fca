#!/bin/bash
# ncu evidence for the v4 step (1 GPU): default bench line, launch list of the bench command, --set full on the tower kernel
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_v4.log 2> gpurun_out/bench_v4.err; echo "== bench: $?"; tail -1 gpurun_out/bench_v4.log | cut -c1-300
B="--games 4096 --sims 6 --steps 1 --warmup 1 --no-cpu-baseline --no-perft --no-kernels"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' -c 900 --csv --log-file gpurun_out/launches_step_v4.csv \
   python bench.py $B > gpurun_out/ncu_launches_v4.log 2>&1; echo "== launch list: $?"
export CRL_NO_GRAPH=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trunk4 -s 4 -c 2 -o gpurun_out/prof_trunk_v4 \
   python bench.py $B > gpurun_out/ncu_trunk4.log 2>&1; echo "== trunk4: $?"
ls -la gpurun_out | tail -8
