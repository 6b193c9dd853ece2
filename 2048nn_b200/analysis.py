"""Analysis harness: the distributions the reference's notebooks are built on --
`notebooks/distribution.py:11-36` (5000 fixed-order games, play_mcts_fixed games per mode: rows
[max tile exponent, score, moves]), `distribution2.ipynb` (histograms, Weibull fits of score / moves, max-tile
percentages) and `compare_models.ipynb` (mean log10(score+1) per model, KS / Welch tests) -- regenerated on the B200
at sample sizes the Python engine cannot reach (16 M fixed-order games take ~30 ms of kernel time).

Games run in the persistent playout kernels; the per-game reduction (max tile, histogram, fixed-point log-score
sum) is b2048_game_summary on the device, so a large sweep moves 20 integers to the host, not the games.
`rows=True` additionally returns the notebooks' per-game array.
"""
import argparse

import numpy as np
import torch

from . import engine
from .board import _raise_if_flag, _session_seed
from .network import blob_for

# quality pins published by the reference
REFERENCE_PINS = {
    "fixed_mean_log_score": (3.3546, 3.3508),   # compare_models.ipynb:171,1463 (5000 games each)
    "shipped_mean_log_score": 4.22628,          # compare_models.ipynb:1612 (5000 policy games, 20200213 c64b3)
    "shipped_frac_2048": 0.240,                 # README.md:30, compare_models.ipynb:1656
}


LOG_FRAC_BITS = 32  # fixed-point scale of acc[16] (b2048_game_summary)


def summarize(acc):
    """acc[20] (tensor from engine.game_summary, or a list of Python ints summed over launches) -> the notebooks'
    headline numbers."""
    a = [int(v) for v in (acc.cpu().tolist() if torch.is_tensor(acc) else acc)]
    games = a[19]
    hist = {int(2 ** e): a[e] for e in range(1, 16) if a[e]}
    return dict(games=games, mean_log_score=a[16] / 2.0 ** LOG_FRAC_BITS / max(games, 1), mean_moves=a[17] / max(games, 1),
                max_score=a[18], maxtile_hist=hist, frac_2048=sum(a[11:16]) / max(games, 1))


def _collect(play, n, chunk, rows):
    total, out = [0] * 20, []
    chunk = min(chunk, 1 << 24)  # one device accumulator per launch; launches are summed in Python integers
    for g0 in range(0, n, chunk):
        m = min(chunk, n - g0)
        final, score, moves = play(g0, m)
        acc, mt = engine.game_summary(final, score, moves, None, want_maxtile=rows)
        part = [int(v) for v in acc.cpu().tolist()]
        total = [max(a, b) if i == 18 else a + b for i, (a, b) in enumerate(zip(total, part))]
        if rows:
            out.append(torch.stack([mt.to(torch.int64), score.to(torch.int64) & 0xFFFFFFFF, moves.to(torch.int64)], 1).cpu().numpy())
    _raise_if_flag()
    s = summarize(total)
    if rows:
        s["rows"] = np.concatenate(out)  # distribution.py:12-14: [max(get_tiles(board)), score, moves]
    return s


def fixed_distribution(n=5000, seed=None, chunk=1 << 24, rows=False, device="cuda"):
    """distribution.py:11-14: n fixed-order (L,U,D,R) games from fresh boards."""
    s = _session_seed(seed)
    return _collect(lambda g0, m: engine.playout_fixed(m, s, g0, device=device), n, chunk, rows)


def policy_distribution(model, n=5000, seed=None, precision=None, chunk=1 << 20, rows=False):
    """compare_models.ipynb's per-model evaluation (eval_nn.py:4-58 with number=n): n policy games."""
    blob, prec = blob_for(model, precision)
    s = _session_seed(seed)

    def play(g0, m):
        r = engine.playout_nn(blob, prec, m, s, g0)
        return r["final"], r["score"], r["moves"]
    return _collect(play, n, chunk, rows)


def mcts_distribution(number, times, mode=0, seed=None, max_steps=8192):
    """distribution.py:17-36: `times` main games of play_mcts_fixed(number, mode) -> rows [max tile, score, moves].
    All games run concurrently in one launch of the device-resident main-game kernel (game g = stream (seed, g))."""
    s = _session_seed(seed)
    while True:
        r = engine.selfplay_fixed_games(int(times), int(number), s, times=1, mode=int(mode), max_steps=max_steps, record=False)
        _raise_if_flag()
        if not bool((r["status"] != 0).any().item()):
            break
        max_steps *= 8
    _, mt = engine.game_summary(r["final"], r["score"].to(torch.int32), r["steps"], want_maxtile=True)
    return torch.stack([mt.to(torch.int64), r["score"], r["steps"].to(torch.int64)], 1).cpu().numpy()


def weibull_fit(x, loc, scale):
    """distribution2.ipynb: weibull_min.fit(dist, loc=0, scale=mean) for scores, (loc=14, scale=median) for moves."""
    from scipy.stats import weibull_min
    return weibull_min.fit(np.asarray(x, dtype=np.float64), loc=loc, scale=scale)


def compare(a, b):
    """compare_models.ipynb stats_test: two-sample KS and Welch t-test."""
    from scipy import stats
    return stats.ks_2samp(a, b), stats.ttest_ind(a, b, equal_var=False)


def main(argv=None):
    p = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    p.add_argument("--fixed", type=int, default=5000, help="fixed-order games")
    p.add_argument("--policy", type=int, default=0, help="policy games of --model")
    p.add_argument("--model", type=str, default="", help="c64b3 checkpoint (.pt state_dict)")
    p.add_argument("--precision", type=str, default="fp16")
    p.add_argument("--seed", type=int, default=12345)
    p.add_argument("--save", type=str, default="", help="write rows to this .npy (distribution/fixed.npy format)")
    a = p.parse_args(argv)
    if a.fixed:
        s = fixed_distribution(a.fixed, a.seed, rows=bool(a.save))
        if a.save:
            np.save(a.save, s.pop("rows"))
        print("fixed-order:", s, "reference pins (5000 games):", REFERENCE_PINS["fixed_mean_log_score"])
    if a.policy:
        from .network import ConvNet
        m = ConvNet(64, 3, precision=a.precision)
        m.load_state_dict(torch.load(a.model, map_location="cpu"))
        s = policy_distribution(m.cuda().eval(), a.policy, a.seed)
        print("policy:", s, "reference pins:", REFERENCE_PINS["shipped_mean_log_score"], REFERENCE_PINS["shipped_frac_2048"])


if __name__ == "__main__":
    main()
