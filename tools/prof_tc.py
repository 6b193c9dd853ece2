"""Short driver for ncu: a few launches of the tensor-core forward and of the NN playout kernel."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("2048nn_b200")
ng = np.load("tests/golden/c64b3_golden.npz")
sd = {k[len("sd_shipped/"):]: torch.from_numpy(np.array(ng[k])) for k in ng.files if k.startswith("sd_shipped/")}
m = pkg.network.ConvNet(64, 3); m.load_state_dict(sd); m.cuda().eval()
boards = torch.from_numpy(ng["lg_boards"].view(np.int64).copy()).cuda().repeat(16)[: 4736 * 24].contiguous()
for _ in range(3):
    m.forward_boards(boards, precision="fp16")
torch.cuda.synchronize()
r = pkg.engine.playout_nn(m.blob("fp16"), "fp16", 4736 * 2, 7, 0)
torch.cuda.synchronize()
print("ok", int(r["total_moves"].item()))
