"""BASELINE config 1: one full main game driven by mcts_fixed (10 playouts per legal move), wall-clock, device-resident
kernel vs the step-by-step host loop; and the batched shapes (selfplay_fixed data generation)."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("2048nn_b200")
mcts = pkg.mcts
mcts.play_mcts_fixed(number=10, seed=1)  # warm-up
mcts._play_mcts_fixed_host(number=10, seed=1)
for name, fn in (("device-resident", mcts.play_mcts_fixed), ("host loop", mcts._play_mcts_fixed_host)):
    for seed in (12345, 2, 3):
        torch.cuda.synchronize(); t = time.perf_counter()
        b, s, c = fn(number=10, seed=seed)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
        print(f"play_mcts_fixed(number=10) {name} seed {seed}: {c} main moves, score {s}, {dt*1e3:.1f} ms  ({dt / c * 1e6:.0f} us per main move)", flush=True)
for number, mode in ((50, 0), (500, 2)):
    torch.cuda.synchronize(); t = time.perf_counter()
    b, s, c = mcts.play_mcts_fixed(number=number, mode=mode, seed=7)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"play_mcts_fixed(number={number}, mode={mode}) device-resident: {c} main moves, score {s}, {dt*1e3:.1f} ms", flush=True)
# batched shapes: concurrent main games, fixed-order min-move search x4 (selfplay_fixed semantics)
for G in (256, 4096):
    pkg.selfplay.selfplay_batch(8, 10, None, 4, seed=5)
    torch.cuda.synchronize(); t = time.perf_counter()
    recs, mv = pkg.selfplay.selfplay_batch(G, 10, None, 4, seed=5)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"selfplay_batch({G} games, fixed10 x4) device-resident: {sum(len(r['boards']) for r in recs)} main moves, {mv} playout moves in {dt:.2f} s "
          f"({mv/dt:.3e} playout moves/s incl. host record assembly), mean score {np.mean([r['score'] for r in recs]):.0f}", flush=True)
t = time.perf_counter()
recs, mv = pkg.selfplay._selfplay_batch_steps(256, 10, None, 4, seed=5)
torch.cuda.synchronize(); dt = time.perf_counter() - t
print(f"selfplay_batch(256 games, fixed10 x4) lock-step host loop: {sum(len(r['boards']) for r in recs)} main moves in {dt:.2f} s")
