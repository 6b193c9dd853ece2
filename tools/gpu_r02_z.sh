#!/bin/bash
# adaptive tail speculation, whole-game launches only: identity test, A/B of eval_nn-style launches and the population launch
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_capi_robust_gpu.py tests/test_population_gpu.py -x -q -m gpu 2>&1 | tail -4
for mode in 1 3; do
echo "== B2048_LATENCY_MODE=$mode"
B2048_LATENCY_MODE=$mode timeout 200 python tools/nn_playout_bench.py 2>&1 | tail -6
B2048_LATENCY_MODE=$mode timeout 200 python tools/population_bench.py 2>&1 | tail -2
done 2>&1 | tee gpurun_out/r02z_tail_speculation.txt
