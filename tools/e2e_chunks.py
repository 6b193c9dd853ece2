"""End-to-end fixed-order playouts through host buffers (engine.HostPlayout): pipeline depth sweep."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("2048nn_b200")
eng = pkg.engine
n = 1 << 24
if len(sys.argv) > 1:
    eng.set_playout_variant(int(sys.argv[1]))
h = torch.empty(n, dtype=torch.int64).pin_memory(); dd = torch.empty(n, dtype=torch.int64, device="cuda")
for name, (a, b_) in {"H2D": (dd, h), "D2H": (h, dd)}.items():
    a.copy_(b_, non_blocking=True); torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(); a.copy_(b_, non_blocking=True); t1.record(); torch.cuda.synchronize()
    print(f"{name} 134 MB pinned: {t0.elapsed_time(t1):.2f} ms = {n * 8 / t0.elapsed_time(t1) / 1e6:.1f} GB/s", flush=True)
start = eng.init_boards(n, 777, 0)
h_start = start.cpu().pin_memory()
out = eng.playout_fixed(n, 1, 0, start=start)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(3):
    eng.playout_fixed(n, 2 + k, 0, start=start, out=out)
e1.record(); torch.cuda.synchronize()
print(f"device-resident: {e0.elapsed_time(e1) / 3:.2f} ms per step", flush=True)
for chunks, ramp in [(c, False) for c in (2, 3, 4, 6, 8, 12, 16)] + [(c, True) for c in (5, 7, 9, 11, 15)]:
    hp = eng.HostPlayout(n, chunks=chunks, ramp=ramp)
    hp(h_start, 3000, 0)
    torch.cuda.synchronize()
    e0.record()
    for k in range(4):
        hp(h_start, 4000 + k, 0)
    e1.record(); torch.cuda.synchronize()
    print(f"chunks={chunks:3d}{' ramp' if ramp else '     '}: {e0.elapsed_time(e1) / 4:.2f} ms per step", flush=True)
    del hp
