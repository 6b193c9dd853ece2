"""evolve.py's population evaluation (50 c64b3 networks x eval_nn(number=100)): one multi-model launch vs the
reference's model-after-model loop over the single-model kernel.  Device-timed."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("2048nn_b200")
ng = np.load("tests/golden/c64b3_golden.npz")
sd = {k[len("sd_shipped/"):]: torch.from_numpy(np.array(ng[k])) for k in ng.files if k.startswith("sd_shipped/")}
base = pkg.network.ConvNet(64, 3, precision="fp16"); base.load_state_dict(sd); base.cuda().eval()
M, number = int(sys.argv[1]) if len(sys.argv) > 1 else 50, int(sys.argv[2]) if len(sys.argv) > 2 else 100
torch.manual_seed(0)
pop = pkg.evolve.NetworkPopulation([base] * 5, mutation_rate=0.002)
nets = [base] + [pop.mutate(base).eval() for _ in range(M - 1)]
blobs = torch.stack([m.blob("fp16") for m in nets]).contiguous()
def timed(fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
    return r, e0.elapsed_time(e1)
r, ms = timed(lambda: pkg.engine.playout_nn_population(blobs, "fp16", number, 5))
mv = int(r["total_moves"].item())
print(f"population launch: {M} nets x {number} games: {ms:.1f} ms, {mv} moves, {mv/ms*1e3:.3e} NN moves/s", flush=True)
def loop():
    return [pkg.engine.playout_nn(m.blob("fp16"), "fp16", number, 5) for m in nets]
rs, ms2 = timed(loop)
mv2 = sum(int(x["total_moves"].item()) for x in rs)
print(f"model-after-model: {ms2:.1f} ms, {mv2} moves, {mv2/ms2*1e3:.3e} NN moves/s  (x{ms2/ms:.1f} slower)", flush=True)
assert mv == mv2
