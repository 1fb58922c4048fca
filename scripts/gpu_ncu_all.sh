#!/bin/bash
# ncu evidence for every kernel class (1 GPU). Graph replay is disabled so kernels appear as plain launches.
mkdir -p gpurun_out
export CRL_NO_GRAPH=1
B="--games 4096 --sims 3 --steps 1 --warmup 1 --no-cpu-baseline --no-perft"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_conv_v2 -s 60 -c 4 -o gpurun_out/prof_conv_v2 python bench.py $B > gpurun_out/ncu1.log 2>&1; echo "conv_v2 $?"
timeout 900 ncu --set full --clock-control none -k regex:'k_(select_expand|reply|finalize|encode_rows|softmax_value)' -s 10 -c 10 -o gpurun_out/prof_tree python bench.py $B > gpurun_out/ncu2.log 2>&1; echo "tree $?"
timeout 900 ncu --set full --clock-control none -k regex:'k_(movegen|make|perft|frontier)' -c 12 -o gpurun_out/prof_rules python scripts/perft_probe.py > gpurun_out/ncu3.log 2>&1; echo "rules $?"
timeout 600 python scripts/perft_probe.py --time > gpurun_out/perft_probe.log 2>&1; echo "perft probe $?"; tail -12 gpurun_out/perft_probe.log
