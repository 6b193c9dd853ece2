#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 150 > gpurun_out/r02h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02h_pytest.log
tail -6 gpurun_out/r02h_pytest.log
timeout 300 python bench.py > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02h_bench.json"))
print("fixed %.4g | nn %.4g (%.3f) | fwd %.4g | sp256 %.4g (%.3f) | game1 %.2f ms | clocks %s" % (d["value"], d["nn_policy"]["value"], d["nn_policy"]["roofline"]["frac"], d["nn_policy"]["forward_only"]["boards_per_s_per_gpu"], d["selfplay256"]["value"], d["selfplay256"]["roofline"]["frac"], d["mcts_nn_game"]["ms_per_main_move"], d["clocks"]))
PY
bash tools/exp_trace.sh 3 > gpurun_out/r02h_trace.txt 2>&1; tail -16 gpurun_out/r02h_trace.txt
