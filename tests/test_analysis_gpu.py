"""GPU: analysis harness (notebooks/distribution.py, distribution2.ipynb, compare_models.ipynb): the device-side
summary equals the same statistics computed on the host from the per-game rows, and the reference's published
quality pins are reproduced within their sampling error."""
import numpy as np
import pytest
import torch

from conftest import state_dict_from_golden

pytestmark = pytest.mark.gpu


def test_summary_equals_host_statistics(pkg, orc):
    s = pkg.analysis.fixed_distribution(20000, seed=3, rows=True, chunk=7000)  # ragged chunks
    rows = s["rows"]
    assert rows.shape == (20000, 3)
    final, score, moves = pkg.engine.playout_fixed(20000, 3, 0)
    fin = final.cpu().numpy().view(np.uint64)
    tiles = np.stack([(fin >> np.uint64(4 * k)) & np.uint64(15) for k in range(16)], 1).max(1)  # max(get_tiles(board))
    assert np.array_equal(rows[:, 0], tiles.astype(np.int64))
    assert np.array_equal(rows[:, 1], score.cpu().numpy().astype(np.int64))
    assert np.array_equal(rows[:, 2], moves.cpu().numpy().astype(np.int64))
    assert s["games"] == 20000 and s["max_score"] == rows[:, 1].max()
    assert s["mean_moves"] == pytest.approx(rows[:, 2].mean(), rel=1e-12)
    assert s["mean_log_score"] == pytest.approx(np.log10(rows[:, 1] + 1).mean(), abs=1e-9)
    assert s["maxtile_hist"] == {int(2 ** e): int(c) for e, c in zip(*np.unique(rows[:, 0], return_counts=True))}
    # first games bit-exact against the CPU oracle
    f, sc, cnt = orc.play_fixed_batch(64, 3)
    assert np.array_equal(fin[:64], f) and np.array_equal(rows[:64, 1], sc.astype(np.int64))
    assert np.array_equal(rows[:64, 2], cnt.astype(np.int64))


def test_fixed_order_pin_at_a_million_games(pkg):
    # compare_models.ipynb:171,1463: mean log10(score+1) of 5000 fixed-order games = 3.3546 / 3.3508 (sigma of one
    # such mean ~0.004); 2^20 games pin it 14x tighter
    s = pkg.analysis.fixed_distribution(1 << 20, seed=12345)
    lo, hi = sorted(pkg.analysis.REFERENCE_PINS["fixed_mean_log_score"])
    assert lo - 0.012 < s["mean_log_score"] < hi + 0.012
    assert s["frac_2048"] == 0.0 and 512 in s["maxtile_hist"]
    assert 200 < s["mean_moves"] < 250
    # 16 M games in one accumulator launch + a second chunk: the fixed-point sums must not wrap
    big = pkg.analysis.fixed_distribution((1 << 24) + 1000, seed=12345)
    assert big["games"] == (1 << 24) + 1000 and abs(big["mean_log_score"] - s["mean_log_score"]) < 2e-3
    fit = pkg.analysis.weibull_fit(pkg.analysis.fixed_distribution(5000, seed=1, rows=True)["rows"][:, 2], loc=14, scale=200)
    assert 1.0 < fit[0] < 4.0  # shape parameter of the moves distribution (distribution2.ipynb)


def test_shipped_policy_pins(pkg, ng):
    m = pkg.network.ConvNet(64, 3, precision="fp16")
    m.load_state_dict(state_dict_from_golden(ng, "shipped"))
    s = pkg.analysis.policy_distribution(m.cuda().eval(), 20000, seed=12345)
    pins = pkg.analysis.REFERENCE_PINS
    # compare_models.ipynb:1612 4.22628 over 5000 games (sigma ~0.005); README.md:30 24.0 % reach 2048 (sigma 0.6 %)
    assert abs(s["mean_log_score"] - pins["shipped_mean_log_score"]) < 0.02
    assert abs(s["frac_2048"] - pins["shipped_frac_2048"]) < 0.02
    a = pkg.analysis.policy_distribution(m, 3000, seed=1, rows=True)["rows"][:, 1]
    b = pkg.analysis.policy_distribution(m, 3000, seed=2, rows=True)["rows"][:, 1]
    ks, _ = pkg.analysis.compare(a, b)
    assert ks.pvalue > 1e-3  # two seeds of one model: same distribution


def test_mcts_distribution_rows(pkg):
    # distribution.py:17-36 shape: one row [max tile exponent, score, moves] per play_mcts_fixed game
    rows = pkg.analysis.mcts_distribution(5, 6, mode=2, seed=21)
    assert rows.shape == (6, 3) and (rows[:, 2] > 50).all() and (rows[:, 0] >= 6).all()
    assert np.array_equal(rows, pkg.analysis.mcts_distribution(5, 6, mode=2, seed=21))  # reproducible
    one = pkg.mcts.play_mcts_fixed(number=5, mode=2, seed=21)  # a one-game session: same kind of game
    assert pkg.analysis.mcts_distribution(5, 1, mode=2, seed=21)[0].tolist() == [max(pkg.board.get_tiles(one[0])), one[1], one[2]]
    better = pkg.analysis.mcts_distribution(5, 24, mode=0, seed=3)[:, 1].mean()
    plain = pkg.analysis.fixed_distribution(4000, seed=3, rows=True)["rows"][:, 1].mean()
    assert better > 2 * plain  # root search beats blind L,U,D,R play by a wide margin (distribution2.ipynb)
