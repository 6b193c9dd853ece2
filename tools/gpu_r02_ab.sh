#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_capi_robust_gpu.py tests/test_network_gpu.py tests/test_selfplay_gpu.py tests/test_maingame_gpu.py -x -q -m gpu 2>&1 | tail -6
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/r02ab_precise32.txt
import importlib, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
pkg = importlib.import_module("2048nn_b200"); eng = pkg.engine
ng = np.load("tests/golden/c64b3_golden.npz"); fx = np.load("tests/golden/c64b3_logits_100k.npz")
m = pkg.network.ConvNet(64, 3, precision="fp16x2")
m.load_state_dict({k[len("sd_shipped/"):]: torch.from_numpy(np.array(ng[k])) for k in ng.files if k.startswith("sd_shipped/")})
m.cuda().eval()
t0 = time.perf_counter(); blob = m.blob("fp16x2"); torch.cuda.synchronize(); print("pack fp16x2 blob (calibration included): %.2f s" % (time.perf_counter() - t0))
b = torch.from_numpy(fx["boards"].view(np.int64).copy()).cuda()
lp = m.forward_boards(b, precision="fp16x2").cpu().numpy()
d = np.abs(lp - fx["lg_shipped"]).max(1)
print("precise mode, 150,371 boards: max |dlogp| %.5f, p99.99 %.5f, >5e-3: %d, argmax agreement %.4f %%" % (d.max(), np.quantile(d, 0.9999), (d >= 5e-3).sum(), 100 * (lp.argmax(1) == fx["lg_shipped"].argmax(1)).mean()))
fb = eng.init_boards(1 << 20, 99, 0)
for prec in ("fp16x2", "fp16"):
    for _ in range(2): m.forward_boards(fb, precision=prec)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): m.forward_boards(fb, precision=prec)
    e1.record(); torch.cuda.synchronize()
    org = eng.init_boards(256, 4242, 0); bl = m.blob(prec)
    eng.mcts_nn_lines(bl, prec, org, 50, 400, 0); torch.cuda.synchronize()
    q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    q0.record(); r = eng.mcts_nn_lines(bl, prec, org, 50, 401, 0); q1.record(); torch.cuda.synchronize()
    eng.selfplay_nn_games(bl, prec, 1, 50, 500, max_steps=4); torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record(); rg = eng.selfplay_nn_games(bl, prec, 1, 50, 501, max_steps=16); g1.record(); torch.cuda.synchronize()
    print(f"{prec}: forward {3 * (1 << 20) / e0.elapsed_time(e1) * 1e3:.4g} boards/s; root evaluation {int(r['total_moves'].item()) / q0.elapsed_time(q1) * 1e3:.4g} moves/s; one main game {g0.elapsed_time(g1) / 16:.2f} ms per main move")
PY
