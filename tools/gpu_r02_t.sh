#!/bin/bash
# full verification of HEAD on one B200: GPU tests, smoke, both bench arms
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/r02t_pytest.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02t_pytest.txt
timeout 200 python __graft_entry__.py smoke > gpurun_out/r02t_smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02t_smoke.txt
( time timeout 400 python bench.py ) > gpurun_out/r02t_bench.json 2> gpurun_out/r02t_bench.err; echo "bench rc=$?"; tail -4 gpurun_out/r02t_bench.err
( time timeout 300 python bench.py --impl reference ) > gpurun_out/r02t_bench_ref.json 2> gpurun_out/r02t_bench_ref.err; echo "ref rc=$?"; tail -4 gpurun_out/r02t_bench_ref.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02t_bench.json").read().strip().splitlines()[-1])
print("fixed %.4g e2e %.4g | nn %.4g (%.3f) %.1f ms | fwd %.4g | sp256 %.4g (%.3f) | game1 %.2f ms | clocks %s" % (d["value"], d["e2e"]["value"], d["nn_policy"]["value"], d["nn_policy"]["roofline"]["frac"], d["nn_policy"]["ms_per_step"], d["nn_policy"]["forward_only"]["boards_per_s_per_gpu"], d["selfplay256"]["value"], d["selfplay256"]["roofline"]["frac"], d["mcts_nn_game"]["ms_per_main_move"], d["clocks"]))
PY
