import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("2048nn_b200")
eng = pkg.engine
eng.set_playout_variant(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
for i in range(3):
    out = eng.playout_fixed(1 << 20, i, 0)
torch.cuda.synchronize()
print("ok", int(out[2].sum().item()))
