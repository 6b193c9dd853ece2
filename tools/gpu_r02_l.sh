#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 150 > gpurun_out/r02l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02l_pytest.log
tail -4 gpurun_out/r02l_pytest.log
timeout 300 python bench.py > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02l_bench.json"))
print("fixed %.4g e2e %.4g %s | nn %.4g (%.3f) | fwd %.4g | sp256 %.4g (%.3f) | game1 %.2f ms | clocks %s" % (d["value"], d["e2e"]["value"], d["e2e"]["pcie_probe"], d["nn_policy"]["value"], d["nn_policy"]["roofline"]["frac"], d["nn_policy"]["forward_only"]["boards_per_s_per_gpu"], d["selfplay256"]["value"], d["selfplay256"]["roofline"]["frac"], d["mcts_nn_game"]["ms_per_main_move"], d["clocks"]))
PY
timeout 120 python tools/prof_k2_full.py > gpurun_out/r02l_plain.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:playout_fixed_kernel3 -s 1 -c 1 \
    -o gpurun_out/r02l_k2v3 python tools/prof_k2_full.py > gpurun_out/r02l_ncu_k2.log 2>&1
echo "ncu k2 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02l_launches.csv \
    python bench.py --steps 2 > gpurun_out/r02l_bench_ncu.log 2>&1
echo "ncu launches rc=$?"
timeout 100 python tools/sweeps.py > gpurun_out/r02l_sweeps.txt 2>&1; tail -30 gpurun_out/r02l_sweeps.txt
