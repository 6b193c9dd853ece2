"""GPU: device-resident main games (b2048_selfplay_fixed) against the step-by-step host loops they replace
(mcts.py:106-148 play_mcts_fixed, selfplay.py:9-59 selfplay_fixed), which are themselves pinned to the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode,number", [(0, 10), (1, 10), (2, 10), (1, 7), (2, 3), (0, 4)])
def test_play_mcts_fixed_device_equals_host_loop(pkg, mode, number):
    for seed in (1, 12345):
        dev = pkg.mcts.play_mcts_fixed(number=number, mode=mode, seed=seed)
        host = pkg.mcts._play_mcts_fixed_host(number=number, mode=mode, seed=seed)
        assert dev == host, (mode, number, seed)
        assert dev[2] > 50 and dev[1] > 0


def test_play_mcts_fixed_from_origin_and_terminal_board(pkg, orc):
    origin = 0x2141348
    assert pkg.mcts.play_mcts_fixed(origin, number=6, mode=2, seed=3) == pkg.mcts._play_mcts_fixed_host(origin, 6, False, 2, 3)
    dead = 0x1212212112122121  # no legal move: the game ends at once (mcts.py:138)
    assert orc.legal_mask(dead) == 0
    assert pkg.mcts.play_mcts_fixed(dead, number=5, mode=0, seed=1) == (dead, 0, 0)


def test_root_values_of_first_step_match_mcts_fixed(pkg):
    # results row 0 of the record = the root values the host functions return for the start board
    for mode, fn in ((0, pkg.mcts.mcts_fixed), (1, pkg.mcts.mcts_fixed_moves), (2, pkg.mcts.mcts_fixed_min)):
        r = pkg.engine.selfplay_fixed_games(1, 10, 77, times=1, mode=mode, max_steps=8)
        start = int(pkg.board._host(r["boards"][0, :1])[0])
        want = fn(start, 10, seed=77, eval_id=1)
        got = r["results"][0, 0].cpu().numpy()
        assert np.allclose(got, np.asarray(want, dtype=np.float32), rtol=0, atol=1e-6), (mode, got, want)
        assert int(r["status"][0]) == 1 and int(r["steps"][0]) == 8  # stopped by the step budget


@pytest.mark.parametrize("times", [1, 3])
def test_selfplay_batch_device_equals_lockstep_loop(pkg, times):
    G, number = 5, 6
    dev, mv_dev = pkg.selfplay.selfplay_batch(G, number, None, times, seed=9)
    host, mv_host = pkg.selfplay._selfplay_batch_steps(G, number, None, times, seed=9)
    assert len(dev) == len(host) == G
    for a, b in zip(dev, host):
        assert np.array_equal(a["boards"], b["boards"])
        assert np.array_equal(a["results"], b["results"])
        assert a["score"] == b["score"]
        assert a["boards"].dtype == np.uint64 and a["results"].dtype == np.float32 and a["results"].shape[1] == 4
    assert mv_dev > 0


def test_shards_of_a_session_play_the_same_games(pkg):
    whole = pkg.engine.selfplay_fixed_games(6, 5, 4, times=2, mode=2, max_steps=2048)
    part = pkg.engine.selfplay_fixed_games(3, 5, 4, times=2, mode=2, game_id_base=3, games_total=6, max_steps=2048)
    assert torch.equal(whole["final"][3:], part["final"]) and torch.equal(whole["score"][3:], part["score"])
    assert torch.equal(whole["steps"][3:], part["steps"])
    n = int(part["steps"][0])
    assert torch.equal(whole["boards"][3, :n], part["boards"][0, :n])
