#!/bin/bash
# three-ply speculation: identity test across latency modes, timing of one main game per mode
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_capi_robust_gpu.py -x -q -m gpu -k latency 2>&1 | tail -15
timeout 200 python tools/debug_maingame.py 2>&1 | grep -v "^G=" | tee gpurun_out/r02x_latency_modes.txt
