"""Short driver for ncu (round 2): one launch each of the tensor-core kernels on their bench workloads, then K2.
TC launch order: [0] warm-up, [1] main games 256 x 2 main moves (two tile-sets), [2] one main game x 8 main moves
(16-warp latency variant), [3] mcts_nn root evaluation 256 origins x 4 x 50 lines, [4] forward 1M boards."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("2048nn_b200")
eng = pkg.engine
ng = np.load("tests/golden/c64b3_golden.npz")
sd = {k[len("sd_shipped/"):]: torch.from_numpy(np.array(ng[k])) for k in ng.files if k.startswith("sd_shipped/")}
m = pkg.network.ConvNet(64, 3, precision="fp16"); m.load_state_dict(sd); m.cuda().eval()
blob = m.blob("fp16")
eng.selfplay_nn_games(blob, "fp16", 1, 4, 5, max_steps=1)
torch.cuda.synchronize()
a = eng.selfplay_nn_games(blob, "fp16", 256, 50, 5, max_steps=2)
b = eng.selfplay_nn_games(blob, "fp16", 1, 50, 5, max_steps=8)
origins = eng.init_boards(256, 4242, 0)
c = eng.mcts_nn_lines(blob, "fp16", origins, 50, 200, 0)
fb = eng.init_boards(1 << 20, 99, 0)
m.forward_boards(fb, precision="fp16")
n = 1 << 24
start = eng.init_boards(n, 777, 0)
f = eng.playout_fixed(n, 2000, 0, start=start)
torch.cuda.synchronize()
print("ok", int(a["total_moves"].item()), int(b["total_moves"].item()), int(c["total_moves"].item()), int(f[2].sum().item()))
