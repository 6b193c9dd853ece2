"""BASELINE config 2 (play_fixed_batch 1K..16M concurrent games) and config 5 (c64b3 forward 1K..1M boards) sweeps,
device-timed (CUDA events, best of 3 after a warm-up).  Two figures per size: one ISOLATED call (event, call, event: the
interval contains the host's own ~0.1 ms of Python / ctypes work before the launch reaches the idle GPU) and the same call
issued 8 times BACK TO BACK (the GPU never idles: what a caller that streams batches sees)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("2048nn_b200")
eng = pkg.engine
def best(fn, reps=3):
    fn(); torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
        out.append((e0.elapsed_time(e1), r))
    return min(out, key=lambda t: t[0])
def streamed(fn, calls=8, reps=3):
    """ms per call with `calls` calls enqueued back to back (best of `reps`)"""
    fn(); torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(2_000_000)  # ~1 ms of spinning on the stream: the calls below queue up behind it
        e0.record()
        for _ in range(calls):
            fn()
        e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / calls)
    return min(out)
print("config 2: play_fixed_batch, fresh boards, fixed L,U,D,R order")
print(f"{'games':>10} {'ms':>9} {'moves/s':>12} {'streamed ms':>12} {'moves/s':>12}")
for k in range(10, 25, 2):
    n = 1 << k
    ms, r = best(lambda: eng.playout_fixed(n, 3, 0))
    ms8 = streamed(lambda: eng.playout_fixed(n, 3, 0), calls=8 if n <= (1 << 20) else 3)
    mv = int(r[2].sum().item())
    print(f"{n:>10} {ms:9.3f} {mv / ms * 1e3:12.3e} {ms8:12.3f} {mv / ms8 * 1e3:12.3e}", flush=True)
print("config 2b: every game from the mid-game origin 0x2141348")
for k in (10, 16, 20, 24):
    n = 1 << k
    start = torch.full((n,), 0x2141348, dtype=torch.int64, device="cuda")
    ms, r = best(lambda: eng.playout_fixed(n, 3, 0, start=start))
    print(f"{n:>10} {ms:9.3f} {int(r[2].sum().item()) / ms * 1e3:12.3e}")
ng = np.load("tests/golden/c64b3_golden.npz")
sd = {k[len("sd_shipped/"):]: torch.from_numpy(np.array(ng[k])) for k in ng.files if k.startswith("sd_shipped/")}
m = pkg.network.ConvNet(64, 3, precision="fp16"); m.load_state_dict(sd); m.cuda().eval()
src = torch.from_numpy(ng["lg_boards"].view(np.int64).copy()).cuda()
print("config 5: c64b3 forward (policy-game positions), packed uint64 boards in, log-probs out")
print(f"{'boards':>10} {'fp16 tcgen05 ms':>16} {'boards/s':>11} {'TFLOP/s valid':>14} {'streamed ms':>12} {'boards/s':>11} {'TFLOP/s valid':>14} {'fp32 exact boards/s':>20}")
for k in range(10, 21, 2):
    n = 1 << k
    b = src.repeat((n + len(src) - 1) // len(src))[:n].contiguous()
    ms, _ = best(lambda: m.forward_boards(b, precision="fp16"))
    ms8 = streamed(lambda: m.forward_boards(b, precision="fp16"))
    ms32, _ = best(lambda: m.forward_boards(b, precision="fp32"), reps=1) if n <= (1 << 16) else (float("nan"), None)
    print(f"{n:>10} {ms:16.3f} {n / ms * 1e3:11.3e} {n / ms * 1e3 * 5128704 / 1e12:14.1f} {ms8:12.3f} {n / ms8 * 1e3:11.3e} "
          f"{n / ms8 * 1e3 * 5128704 / 1e12:14.1f} {n / ms32 * 1e3:20.3e}", flush=True)
