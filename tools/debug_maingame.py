"""Bring-up timing of the device-resident NN main-game kernel at growing sizes (prints as it goes)."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("2048nn_b200")
eng = pkg.engine
ng = np.load("tests/golden/c64b3_golden.npz")
tag = sys.argv[1] if len(sys.argv) > 1 else "shipped"
m = pkg.network.ConvNet(64, 3, precision="fp16")
m.load_state_dict({k[len(f"sd_{tag}/"):]: torch.from_numpy(np.array(ng[k])) for k in ng.files if k.startswith(f"sd_{tag}/")})
m.cuda().eval()
blob = m.blob("fp16")
eng.selfplay_nn_games(blob, "fp16", 1, 4, 5, max_steps=1)
torch.cuda.synchronize()
ref = {}
for wide in (3, 2, 1, 0):
    eng.set_latency_mode(wide)
    for G, number, steps in ((1, 50, 16), (4, 50, 8), (8, 50, 4)):
        t0 = time.perf_counter()
        r = eng.selfplay_nn_games(blob, "fp16", G, number, 5, max_steps=steps)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        mv = int(r["total_moves"].item()); st = int(r["steps"].sum().item())
        key = (G, number, steps)
        same = "" if key not in ref else (" same games as latency mode: %s" % bool(torch.equal(ref[key], r["boards"])))
        ref.setdefault(key, r["boards"].clone())
        print(f"latency_mode={wide} G={G} number={number} max_steps={steps}: {dt*1e3:.1f} ms, {st} main moves, {mv} policy moves, {mv/dt:.3e} moves/s, "
              f"{dt*1e3/max(1,st)*G:.2f} ms per main move per game{same}", flush=True)
    b = torch.from_numpy(ng["lg_boards"][:1024].view(np.int64).copy()).cuda()
    m.forward_boards(b); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        lp = m.forward_boards(b)
    e1.record(); torch.cuda.synchronize()
    print(f"latency_mode={wide} forward 1024 boards: {e0.elapsed_time(e1)/20*1e3:.1f} us, max|dlogp| vs reference {np.abs(lp.cpu().numpy() - ng['lg_shipped'][:1024]).max():.4f}", flush=True)
eng.set_latency_mode(3)
for G, number, steps in ((64, 50, 4), (256, 50, 2), (256, 50, 8), (32, 50, 8)):
    t0 = time.perf_counter()
    r = eng.selfplay_nn_games(blob, "fp16", G, number, 5, max_steps=steps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    mv = int(r["total_moves"].item()); st = int(r["steps"].sum().item())
    print(f"G={G} number={number} max_steps={steps}: {dt*1e3:.1f} ms, {st} main moves, {mv} policy moves, {mv/dt:.3e} moves/s, "
          f"{dt*1e3/max(1,st)*G:.2f} ms per main move per game, err flag {pkg._lib.take_error_flag()}", flush=True)
