#!/bin/bash
# ncu --set full on the device-driven perft kernels (crl_perft_root_host): breadth-first ply + per-lane walk
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_(bfs_ply|perft_walk)' -c 40 -o gpurun_out/prof_perft_root \
   python scripts/perft_root_probe.py > gpurun_out/ncu_perft_root.log 2>&1; echo "== ncu perft_root: $?"
