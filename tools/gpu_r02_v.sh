#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_dropin_gpu.py -x -q -m gpu 2>&1 | tail -5
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
for lanes in 2 3 2 3; do
timeout 200 python bench.py --no-train --nn-lanes $lanes > gpurun_out/r02v_bench_l$lanes.json 2> gpurun_out/r02v_bench_l$lanes.err; echo "bench lanes=$lanes rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02v_bench_l$lanes.json").read().strip().splitlines()[-1])
print("lanes $lanes: nn %.4g (%.3f) %.2f ms/step" % (d["nn_policy"]["value"], d["nn_policy"]["roofline"]["frac"], d["nn_policy"]["ms_per_step"]))
PY
done
