#!/bin/bash
mkdir -p gpurun_out
echo "--- default"; timeout 200 python scripts/probe/reuse_debug.py 2>&1 | cut -c1-1500 | tail -60
