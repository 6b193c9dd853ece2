#!/bin/bash
# Build the -DB2048_DEBUG_ASSERTS variant next to the product library and run the tensor-core workloads on it
# (forward, playouts in every latency mode, main games incl. the queueing regime, precise mode).  A violated assert traps.
set -e
export B2048_LIB=$PWD/2048nn_b200/lib2048nn_b200_dbg.so
B2048_EXP=B2048_DEBUG_ASSERTS python 2048nn_b200/build.py > /dev/null
python - <<'PY'
import importlib, numpy as np, torch
pkg = importlib.import_module("2048nn_b200")
assert pkg.LIB_PATH.endswith("_dbg.so")
eng = pkg.engine
ng = np.load("tests/golden/c64b3_golden.npz")
m = pkg.network.ConvNet(64, 3, precision="fp16")
m.load_state_dict({k[len("sd_random/"):]: torch.from_numpy(np.array(ng[k])) for k in ng.files if k.startswith("sd_random/")})
m.cuda().eval()
b = torch.from_numpy(ng["lg_boards"].view(np.int64).copy()).cuda()
for mode in (0, 1, 2, 3):
    eng.set_latency_mode(mode)
    for prec in ("fp16", "fp16x2"):
        blob = m.blob(prec)
        lp = m.forward_boards(b, precision=prec)
        lp2 = m.forward_boards(b[:700].contiguous(), precision=prec)
        r = eng.playout_nn(blob, prec, 300, 5, 0)
        r2 = eng.playout_nn(blob, prec, 12000, 6, 0, group_size=40)
        o = eng.init_boards(40, 7, 0)
        r3 = eng.mcts_nn_lines(blob, prec, o, 50, 8, 0)
        r4 = eng.selfplay_nn_games(blob, prec, 2, 10, 9, max_steps=64)
        r5 = eng.selfplay_nn_games(blob, prec, 70, 40, 10, max_steps=6)
        torch.cuda.synchronize()
        assert np.abs(lp.cpu().numpy() - ng["lg_random"]).max() < 1e-3
        print("mode", mode, prec, "ok:", int(r["total_moves"].item()), int(r2["total_moves"].item()), int(r3["total_moves"].item()),
              int(r4["steps"].sum().item()), int(r5["total_moves"].item()), flush=True)
print("debug-assert build: all workloads ran without a trap")
PY
