#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_capi_robust_gpu.py tests/test_network_gpu.py -x -q -m gpu -k "latency or precise" 2>&1 | tail -4
timeout 200 python - <<'PY' 2>&1 | tee gpurun_out/r02aa_precise_latency.txt
import importlib, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
pkg = importlib.import_module("2048nn_b200"); eng = pkg.engine
ng = np.load("tests/golden/c64b3_golden.npz")
m = pkg.network.ConvNet(64, 3, precision="fp16x2")
m.load_state_dict({k[len("sd_shipped/"):]: torch.from_numpy(np.array(ng[k])) for k in ng.files if k.startswith("sd_shipped/")})
m.cuda().eval()
for prec in ("fp16x2", "fp16"):
    blob = m.blob(prec)
    for mode in (3, 1):
        eng.set_latency_mode(mode)
        eng.selfplay_nn_games(blob, prec, 1, 50, 5, max_steps=2); torch.cuda.synchronize()
        t0 = time.perf_counter(); r = eng.selfplay_nn_games(blob, prec, 1, 50, 5, max_steps=16); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.playout_nn(blob, prec, 4736, 7, 0); torch.cuda.synchronize()
        e0.record(); g = eng.playout_nn(blob, prec, 4736, 7, 0); e1.record(); torch.cuda.synchronize()
        print(f"{prec} latency_mode={mode}: one main game {dt*1e3/16:.2f} ms per mcts_nn(number=50) main move; 4,736 policy games to the end {e0.elapsed_time(e1):.1f} ms ({int(g['total_moves'].item())/e0.elapsed_time(e1)*1e3:.3e} moves/s)", flush=True)
eng.set_latency_mode(3)
PY
