"""One launch of the bench's nn_policy workload (mcts_nn: 256 origins x 4 x 50 lines) for an ncu traffic capture."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("2048nn_b200")
ng = np.load("tests/golden/c64b3_golden.npz")
sd = {k[len("sd_shipped/"):]: torch.from_numpy(np.array(ng[k])) for k in ng.files if k.startswith("sd_shipped/")}
m = pkg.network.ConvNet(64, 3, precision="fp16"); m.load_state_dict(sd); m.cuda().eval()
origins = pkg.engine.init_boards(256, 4242, 0)
r = pkg.dist.mcts_nn_sharded(m.blob("fp16"), "fp16", origins, 50, 200, 0)
torch.cuda.synchronize()
print("ok", int(r["total_moves"].item()))
