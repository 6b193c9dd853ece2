"""BASELINE config 1: one full main game driven by mcts_fixed (10 playouts per legal move), wall-clock."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("2048nn_b200")
mcts = pkg.mcts
mcts.play_mcts_fixed(number=10, seed=1)  # warm-up
for seed in (12345, 2, 3):
    torch.cuda.synchronize(); t = time.perf_counter()
    b, s, c = mcts.play_mcts_fixed(number=10, seed=seed)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"play_mcts_fixed(number=10) seed {seed}: {c} main moves, score {s}, {dt:.3f} s  ({dt / c * 1e6:.0f} us per main move)", flush=True)
# batched shape: 256 concurrent main games, fixed-order min-move search x4 (selfplay_fixed semantics)
t = time.perf_counter()
recs, _ = pkg.selfplay.selfplay_batch(256, 10, None, 4, seed=5)
torch.cuda.synchronize(); dt = time.perf_counter() - t
print(f"selfplay_batch(256 games, fixed10 x4): {sum(len(r['boards']) for r in recs)} main moves in {dt:.2f} s, mean score {np.mean([r['score'] for r in recs]):.0f}")
