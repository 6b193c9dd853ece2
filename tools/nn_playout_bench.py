"""Device-timed NN-policy playout throughput (tensor-core path)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("2048nn_b200")
ng = np.load("tests/golden/c64b3_golden.npz")
tag = "shipped"
sd = {k[len(f"sd_{tag}/"):]: torch.from_numpy(np.array(ng[k])) for k in ng.files if k.startswith(f"sd_{tag}/")}
m = pkg.network.ConvNet(64, 3); m.load_state_dict(sd); m.cuda().eval()
blob = m.blob("fp16")
def timed(fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
    return r, e0.elapsed_time(e1)
for n in (4736, 4736 * 4, 4736 * 16):
    r, ms = timed(lambda: pkg.engine.playout_nn(blob, "fp16", n, 7, 0))
    mv = int(r["total_moves"].item())
    print(f"eval_nn-style {n} games: {ms:.1f} ms, {mv} moves, {mv/ms*1e3:.3e} NN moves/s, mean len {mv/n:.0f}", flush=True)
origins = pkg.engine.init_boards(256, 3, 0)
for number in (50,):
    r, ms = timed(lambda: pkg.engine.mcts_nn_lines(blob, "fp16", origins, number, 9, 0))
    mv = int(r["total_moves"].item())
    print(f"mcts_nn C4-style 256 origins x 4 x {number}: {ms:.1f} ms, {mv} moves, {mv/ms*1e3:.3e} NN moves/s", flush=True)
o1 = origins[:1].contiguous()
r, ms = timed(lambda: pkg.engine.mcts_nn_lines(blob, "fp16", o1, 50, 9, 0))
print(f"mcts_nn C3 single origin x 4 x 50: {ms:.2f} ms, {int(r['total_moves'].item())} moves", flush=True)
