#!/bin/bash
# Runs on the GPU box under gpurun: GPU parity tests (one process per file so a CUDA fault in one file does not
# poison the rest), smoke, and a short bench.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for f in rules encode tree net; do
  timeout 600 python -m pytest tests/test_gpu_$f.py -m gpu -q -x --timeout 500 > gpurun_out/pytest_$f.log 2>&1
  echo "== $f: exit $?"; tail -5 gpurun_out/pytest_$f.log
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke: exit $?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py ${BENCH_ARGS:---games 512 --sims 32 --steps 2 --warmup 1 --no-cpu-baseline} > gpurun_out/bench_small.log 2>&1; echo "== bench: exit $?"; tail -3 gpurun_out/bench_small.log
