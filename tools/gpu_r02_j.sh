#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 150 > gpurun_out/r02j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02j_pytest.log
tail -6 gpurun_out/r02j_pytest.log
timeout 300 python bench.py > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02j_bench.json"))
print("fixed %.4g e2e %.4g | nn %.4g (%.3f) | fwd %.4g | sp256 %.4g (%.3f) | game1 %.2f ms | clocks %s" % (d["value"], d["e2e"]["value"], d["nn_policy"]["value"], d["nn_policy"]["roofline"]["frac"], d["nn_policy"]["forward_only"]["boards_per_s_per_gpu"], d["selfplay256"]["value"], d["selfplay256"]["roofline"]["frac"], d["mcts_nn_game"]["ms_per_main_move"], d["clocks"]))
PY
timeout 120 python tools/prof_k2_full.py > gpurun_out/r02j_plain.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:playout_fixed_kernel3 -s 1 -c 1 \
    -o gpurun_out/r02j_k2v3 python tools/prof_k2_full.py > gpurun_out/r02j_ncu_k2.log 2>&1
echo "ncu k2 rc=$?"
timeout 100 python tools/quick_k2.py 2
