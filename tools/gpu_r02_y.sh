#!/bin/bash
# adaptive tail speculation: identity tests, then A/B of the nn_policy block (latency mode 1 = plain kernel, 3 = adaptive)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_capi_robust_gpu.py tests/test_network_gpu.py tests/test_population_gpu.py tests/test_selfplay_gpu.py -x -q -m gpu 2>&1 | tail -6
for mode in 1 3 1 3; do
B2048_LATENCY_MODE=$mode timeout 200 python bench.py --no-train > gpurun_out/r02y_bench_m$mode.json 2> gpurun_out/r02y_bench_m$mode.err; echo "bench mode=$mode rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02y_bench_m$mode.json").read().strip().splitlines()[-1])
print("mode $mode: nn %.4g (%.3f) %.2f ms/step e2e %.4g | game1 %.2f ms" % (d["nn_policy"]["value"], d["nn_policy"]["roofline"]["frac"], d["nn_policy"]["ms_per_step"], d["nn_policy"]["e2e"]["value"], d["mcts_nn_game"]["ms_per_main_move"]))
PY
done
timeout 200 python tools/population_bench.py 2>&1 | tail -4
