"""Per-step and per-chunk timing of engine.HostPlayout (e2e path of bench.py): where does a slow step lose its time?"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
if "LOCAL_RANK" in os.environ:
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
pkg = importlib.import_module("2048nn_b200")
eng = pkg.engine
n = 1 << 24
start = eng.init_boards(n, 777, 0)
h_start = start.cpu().pin_memory()
hp = eng.HostPlayout(n)
hp(h_start, 3000, 0)
torch.cuda.synchronize()
rank = os.environ.get("RANK", "0")
for k in range(12):
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    hp(h_start, 4000 + k, 0)
    e1.record(); torch.cuda.synchronize()
    t1 = time.perf_counter()
    ev = getattr(hp, "last_events", None)
    extra = ""
    if ev:
        base = ev[0][0]
        extra = " | " + " ".join("c%d[in %.1f-%.1f run %.1f-%.1f out -%.1f]" % (i, base.elapsed_time(a), base.elapsed_time(b), base.elapsed_time(c), base.elapsed_time(d_), base.elapsed_time(e))
                                 for i, (a, b, c, d_, e) in enumerate(ev))
    print(f"rank {rank} step {k}: device {e0.elapsed_time(e1):.2f} ms, host {1e3 * (t1 - t0):.2f} ms{extra}", flush=True)
