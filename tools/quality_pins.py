"""The reference's published quality pins regenerated at large sample sizes on the B200."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("2048nn_b200")
A = pkg.analysis
t = time.time()
s = A.fixed_distribution(1 << 24, seed=12345)
n = s["games"]
print(f"fixed-order L,U,D,R: {n} games in {time.time()-t:.2f} s: mean log10(score+1) = {s['mean_log_score']:.5f} "
      f"(reference, 5000 games each: {A.REFERENCE_PINS['fixed_mean_log_score']}), mean moves {s['mean_moves']:.2f}, max score {s['max_score']}")
print("  max tile histogram:", {k: f"{v / n * 100:.3f}%" for k, v in sorted(s["maxtile_hist"].items())})
print("  reference max-tile fractions over 5000 games (distribution2.ipynb:998-999): 512: 4.44 %, 256: 47.04 %, 128: 38.8 %, 64: 9.16 %")
rows = A.fixed_distribution(200000, seed=7, rows=True)["rows"]
fs = A.weibull_fit(rows[:, 1], loc=0, scale=rows[:, 1].mean())
fm = A.weibull_fit(rows[:, 2], loc=14, scale=np.median(rows[:, 2]))
print(f"  Weibull fits on 200,000 games (distribution2.ipynb: weibull_min.fit): score (c, loc, scale) = ({fs[0]:.3f}, {fs[1]:.1f}, {fs[2]:.1f}) "
      f"[reference (2.064, 176.2, 2685.9)], moves = ({fm[0]:.3f}, {fm[1]:.1f}, {fm[2]:.1f}) [reference (2.457, 43.2, 205.9)]")
ng = np.load("tests/golden/c64b3_golden.npz")
sd = {k[len("sd_shipped/"):]: torch.from_numpy(np.array(ng[k])) for k in ng.files if k.startswith("sd_shipped/")}
for prec in ("fp16", "fp32"):
    m = pkg.network.ConvNet(64, 3, precision=prec); m.load_state_dict(sd); m.cuda().eval()
    N = (1 << 18) if prec == "fp16" else (1 << 14)
    t = time.time()
    p = A.policy_distribution(m, N, seed=12345)
    print(f"shipped c64b3 policy games ({prec}): {N} games in {time.time()-t:.2f} s: mean log10(score+1) = {p['mean_log_score']:.5f} "
          f"(reference, 5000 games: {A.REFERENCE_PINS['shipped_mean_log_score']}), reached 2048: {p['frac_2048'] * 100:.2f} % "
          f"(reference: {A.REFERENCE_PINS['shipped_frac_2048'] * 100:.1f} %), mean moves {p['mean_moves']:.1f}, max score {p['max_score']}")
    print("  max tile histogram:", {k: f"{v / N * 100:.2f}%" for k, v in sorted(p["maxtile_hist"].items())})
