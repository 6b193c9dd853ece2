"""Quick device-timed throughput probe of the fixed-order playout kernel variants (not the bench)."""
import importlib
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

pkg = importlib.import_module("2048nn_b200")
eng = pkg.engine
for variant in (0, 1):
    eng.set_playout_variant(variant)
    for n in [1 << 20, 1 << 24]:
        out = eng.playout_fixed(n, 1, 0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = eng.playout_fixed(n, 2, 0, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        moves = int(out[2].sum().item())
        print(f"variant={variant} n={n} ms={ms:.2f} moves/s={moves/ms*1e3:.3e}", flush=True)
