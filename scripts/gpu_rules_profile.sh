#!/bin/bash
# rule kernels: timing probe, ncu --set full with source for movegen / perft, and the launch list of one bench step
mkdir -p gpurun_out
timeout 600 python scripts/perft_probe.py --time > gpurun_out/perft_probe.log 2>&1; echo "probe $?"; cat gpurun_out/perft_probe.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_(movegen|perft)' -c 8 -o gpurun_out/prof_rules_src python scripts/perft_probe.py > gpurun_out/ncu_rules.log 2>&1; echo "ncu rules $?"
B="--games 4096 --sims 6 --steps 1 --warmup 1 --no-cpu-baseline --no-perft --no-kernels"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' -c 900 --csv --log-file gpurun_out/launches_step_v3.csv \
   python bench.py $B > gpurun_out/ncu_launches_step.log 2>&1; echo "launch list $?"; tail -3 gpurun_out/ncu_launches_step.log | cut -c1-300
