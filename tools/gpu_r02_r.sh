#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 150 > gpurun_out/r02r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02r_pytest.log
tail -4 gpurun_out/r02r_pytest.log
timeout 300 python bench.py > gpurun_out/r02r_bench.json 2> gpurun_out/r02r_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02r_bench.json"))
print("fixed %.4g e2e %.4g %s | nn %.4g (%.3f) | fwd %.4g | sp256 %.4g (%.3f) | game1 %.2f ms | clocks %s" % (d["value"], d["e2e"]["value"], d["e2e"]["pcie_probe"], d["nn_policy"]["value"], d["nn_policy"]["roofline"]["frac"], d["nn_policy"]["forward_only"]["boards_per_s_per_gpu"], d["selfplay256"]["value"], d["selfplay256"]["roofline"]["frac"], d["mcts_nn_game"]["ms_per_main_move"], d["clocks"]))
PY
timeout 300 python bench.py --impl reference > gpurun_out/r02r_bench_ref.json 2> gpurun_out/r02r_bench_ref.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/r02r_bench_ref.json
timeout 200 python __graft_entry__.py smoke > gpurun_out/r02r_smoke.txt 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02r_smoke.txt
