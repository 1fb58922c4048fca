#!/bin/bash
# k_trunk4 single-tile mode for small batches: network parity tests (n in {1, 5, 37, 592, 593, 4096} cover both modes),
# the whole GPU suite, then the small-lane measurements again
mkdir -p gpurun_out
T0=$SECONDS
timeout 600 python -m pytest tests/test_gpu_net.py tests/test_gpu_parity_net.py -q -x --timeout 300 > gpurun_out/pytest_net.log 2>&1; echo "== net tests: $? at $((SECONDS-T0)) s"; tail -3 gpurun_out/pytest_net.log
timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "== pytest -m gpu: $? at $((SECONDS-T0)) s"; tail -3 gpurun_out/pytest_gpu.log
bash scripts/gpu_r2_single_game.sh
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-perft --no-kernels --no-whole-games --no-large --no-training > gpurun_out/b2.log 2> gpurun_out/b2.err
python -c "
import json; d=json.loads(open('gpurun_out/b2.log').read().strip().splitlines()[-1]); print('4096x200', round(d['value']), round(d['e2e']['value']), d['clocks'])"
