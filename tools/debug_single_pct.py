"""A/B of the one-set / two-set choice for main-game launches (B2048_MAIN_SINGLE_PCT is read at context creation)."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("2048nn_b200")
eng = pkg.engine
ng = np.load("tests/golden/c64b3_golden.npz")
m = pkg.network.ConvNet(64, 3, precision="fp16")
m.load_state_dict({k[len("sd_shipped/"):]: torch.from_numpy(np.array(ng[k])) for k in ng.files if k.startswith("sd_shipped/")})
m.cuda().eval()
blob = m.blob("fp16")
eng.selfplay_nn_games(blob, "fp16", 1, 4, 5, max_steps=1)
torch.cuda.synchronize()
for G in (16, 24, 32, 40, 48, 64):
    best = None
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = eng.selfplay_nn_games(blob, "fp16", G, 50, 5, max_steps=8)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    mv = int(r["total_moves"].item())
    print(f"pct={os.environ.get('B2048_MAIN_SINGLE_PCT', 'default')} G={G} ({G * 200} lines): {best:.1f} ms, {mv / best * 1e3:.3e} moves/s", flush=True)
