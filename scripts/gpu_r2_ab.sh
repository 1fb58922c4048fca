#!/bin/bash
for rnd in 1 2; do for v in prev b; do timeout 200 python scripts/trunk4_ab_probe.py gpurun_ab/$v 2>&1 | grep "n="; done; done
timeout 300 python -m pytest tests/test_gpu_net.py tests/test_gpu_parity_net.py -q -x --timeout 200 2>&1 | tail -2
