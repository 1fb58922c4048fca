#!/bin/bash
# ncu --set full of the non-tower step kernels on the fused-head build (CRL_NO_GRAPH=1 so that every launch is visible)
mkdir -p gpurun_out
S="--no-cpu-baseline --no-perft --no-kernels --no-whole-games --no-large"
B="--games 4096 --sims 6 --steps 1 --warmup 1 $S"
CRL_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_(select_expand|reply|finalize|encode_rows|softmax_value|conv_v2)$' -s 18 -c 18 -o gpurun_out/prof_step_small_r02b \
   python bench.py $B > gpurun_out/ncu_step_small_r02b.log 2>&1; echo "== step kernels full: $?"
ls -la gpurun_out | tail -3
