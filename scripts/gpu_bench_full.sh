#!/bin/bash
# full-size bench + ncu launch list + ncu --set full on the conv kernel (1 GPU)
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; echo "== bench full: exit $?"; tail -c 3000 gpurun_out/bench_full.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
   python bench.py --games 4096 --sims 4 --steps 1 --warmup 1 --no-cpu-baseline --no-perft > gpurun_out/ncu_launches.log 2>&1; echo "== ncu launches: exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_conv3x3 -s 30 -c 3 -o gpurun_out/prof_conv \
   python bench.py --games 4096 --sims 2 --steps 1 --warmup 1 --no-cpu-baseline --no-perft > gpurun_out/ncu_conv.log 2>&1; echo "== ncu conv: exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_heads -s 2 -c 1 -o gpurun_out/prof_heads \
   python bench.py --games 4096 --sims 2 --steps 1 --warmup 1 --no-cpu-baseline --no-perft > gpurun_out/ncu_heads.log 2>&1; echo "== ncu heads: exit $?"
ls -la gpurun_out
