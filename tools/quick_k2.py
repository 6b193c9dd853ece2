"""Quick device-timed throughput probe of the fixed-order playout kernel variants (not the bench)."""
import importlib
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

pkg = importlib.import_module("2048nn_b200")
eng = pkg.engine
variants = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else (0, 1, 2)
for variant in variants:
    eng.set_playout_variant(variant)
    for n in [1 << 10, 1 << 16, 1 << 20, 1 << 24]:
        out = eng.playout_fixed(n, 1, 0)
        torch.cuda.synchronize()
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = eng.playout_fixed(n, 2, 0, out=out)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        moves = int(out[2].sum().item())
        print(f"variant={variant} n={n} ms={best:.3f} moves/s={moves/best*1e3:.3e}", flush=True)
