"""ncu driver: the bench's main workload (16M games from resident start boards), 3 launches."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("2048nn_b200")
eng = pkg.engine
n = 1 << 24
start = eng.init_boards(n, 777, 0)
out = None
for i in range(3):
    out = eng.playout_fixed(n, 1000 + i, 0, start=start, out=out)
torch.cuda.synchronize()
print("ok", int(out[2].sum().item()))
