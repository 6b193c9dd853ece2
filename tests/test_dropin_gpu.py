"""GPU: the reference-facing Python entry points (board / mcts drop-ins) against the golden
vectors made by the reference itself and against the oracle.  Reads like the reference's API."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def board(pkg):
    return pkg.board


@pytest.fixture(scope="module")
def mcts(pkg):
    return pkg.mcts


def test_scalar_board_functions(board, eg):
    assert board.transpose(0x0123456789ABCDEF) == 0x048C159D26AE37BF
    assert board.countzero(0) == 0 and board.countzero(0x2141348) == 9
    assert board.get_tiles(0x21)[:3] == [1, 2, 0]
    assert board.move(0x2141348, 2) == (0x21401348, 0, True)
    assert board.move(0x2141348, 0) == (0x2141348, 0, False)
    assert board.move(0xFFFFFFFFFFFFFFFF, 3) == (0, 524288, True)
    r = board.move(0x1111222233334444, 0)
    assert r == (0x0022003300440055, 120, True) and type(r[0]) is int and type(r[2]) is bool
    assert board.merge_table[0x1111] == (0x0022, 8, True) and board.merge_table_rev[0x1111] == (0x2200, 8, True)
    assert board.move_row(0xFFFF, False) == (0, 131072, True)
    for j in range(0, 3000, 97):
        b = int(eg["mv_boards"][j])
        for d in range(4):
            assert board.move(b, d) == (int(eg["mv_out"][d, j]), int(eg["mv_score"][d, j]), bool(eg["mv_moved"][d, j]))
    with pytest.raises(ValueError):
        board.generate_tile(0)  # randrange(0) in the reference (16 empties wrap)


def test_play_fixed_replays_reference_games(board, eg):
    seed = int(eg["pf_seed"])
    for g in (0, 1, 2, 77, 1499):
        assert board.generate_init_tiles(seed, g) == int(eg["pf_start"][g])
        assert board.play_fixed(seed=seed, gid=g) == (int(eg["pf_final"][g]), int(eg["pf_score"][g]), int(eg["pf_count"][g]))
    o, s = int(eg["po_origin"]), int(eg["po_seed"])
    for g in (0, 5, 499):
        assert board.play_fixed(o, seed=s, gid=g) == (int(eg["po_final"][g]), int(eg["po_score"][g]), int(eg["po_count"][g]))


def test_play_fixed_batch_death_order(board, eg):
    n = 1500
    scores = board.play_fixed_batch(n, seed=int(eg["pf_seed"]))
    order = np.argsort(eg["pf_count"], kind="stable")  # by step, then list position (board.py:312-314)
    assert scores == [int(x) for x in eg["pf_score"][order]]
    assert board.play_fixed_batch(0) == []


def test_boardarray_loop_equals_play_fixed(board, eg):
    seed = int(eg["pf_seed"])
    k = 40
    arr = board.BoardArray([int(x) for x in eg["pf_start"][:k]], seed=seed)
    assert arr.scores == [0] * k
    dead = {}
    step = 0
    alive = list(range(k))
    while arr.boards:
        ds, db = arr.move_batch([(0, 1, 3, 2)] * len(arr.boards), get_boards=True)
        died = [g for g in alive if True][:0]
        # survivors stay in order: reconstruct who died from counts
        want_dead = [g for g in alive if int(eg["pf_count"][g]) == step]
        assert ds == [int(eg["pf_score"][g]) for g in want_dead]
        assert db == [int(eg["pf_final"][g]) for g in want_dead]
        alive = [g for g in alive if g not in want_dead]
        step += 1
    # fast_move_batch: True at the first death, array untouched
    arr = board.BoardArray([int(eg["pf_final"][0]), int(eg["pf_start"][1])], seed=seed)
    before = list(arr.boards)
    assert arr.fast_move_batch([(0, 1, 3, 2)] * 2) is True and arr.boards == before


def test_boardarray_policy_rows_vs_oracle(board, orc, eg):
    """move_batch / fast_move_batch with PER-GAME priority rows (the (N,4) int64 tensor eval_nn.py:41-42 and
    mcts_nn.py:36-37 pass) against the oracle's step, over whole games: survivors stay in order, the dead are returned in
    list order with their last boards, scores accumulate from 0; fast_move_batch advances exactly like move_batch while
    every game moves and returns True, array untouched, at the first step with a death (board.py:279-303)."""
    import torch

    seed, n = 97531, 300
    rng = np.random.default_rng(5)
    starts = eg["pf_start"][:n].astype(np.uint64)  # fresh boards: the first death comes after a few dozen steps
    arr = board.BoardArray([int(x) for x in starts], seed=seed)
    fast = board.BoardArray([int(x) for x in starts], seed=seed)
    fast_live = True
    ob, osc = starts.copy(), np.zeros(n, np.uint64)
    sidx, alive = np.zeros(n, np.uint32), np.ones(n, np.uint8)
    keys = np.array([orc.rng_key(seed, g) for g in range(n)], np.uint64)
    steps = 0
    while arr.boards:
        live = np.nonzero(alive)[0]
        rows = np.stack([rng.permutation(4) for _ in live]).astype(np.int64)
        prio = np.zeros((n, 4), np.uint8)
        prio[live] = rows
        before_b, before_s = ob.copy(), osc.copy()
        deaths = orc.policy_step(ob, osc, sidx, keys, prio, alive)
        died = [g for g in live if not alive[g]]
        assert deaths == len(died)
        move_list = torch.from_numpy(rows) if steps % 2 == 0 else [tuple(int(v) for v in r) for r in rows]
        if fast_live:
            snapshot = (list(fast.boards), list(fast.scores))
            hit = fast.fast_move_batch(move_list)
            assert hit is (deaths > 0)
            if hit:
                assert (fast.boards, fast.scores) == snapshot  # untouched
                fast_live = False
        ds, db = arr.move_batch(move_list, get_boards=True)
        assert ds == [int(before_s[g]) for g in died] and db == [int(before_b[g]) for g in died]
        keep = np.nonzero(alive)[0]
        assert arr.boards == [int(ob[g]) for g in keep] and arr.scores == [int(osc[g]) for g in keep]
        if fast_live:
            assert fast.boards == arr.boards and fast.scores == arr.scores
        steps += 1
    assert steps > 50 and not fast_live and len(fast.boards) == n
    assert arr.move_batch([], get_boards=True) == ([], []) and arr.fast_move_batch([]) is False


def test_mcts_fixed_family(mcts, eg):
    number, seed = int(eg["mc_number"]), int(eg["mc_seed"])
    for e, o in enumerate(eg["mc_origins"].tolist()):
        got = mcts.mcts_fixed(o, number, seed=seed, eval_id=e)
        for i in range(4):
            if eg["mc_min"][e, i] == 0:
                assert got[i] == -1
            else:
                order = np.argsort(eg["mc_moves"][e, i], kind="stable")
                s = int(mcts.move(o, i)[1])
                want = np.mean(np.log10(eg["mc_scores"][e, i][order] + s + 1))
                assert got[i] == want and isinstance(got[i], np.float64)
                assert abs(got[i] - eg["mc_mean"][e, i]) < 1e-12
        assert mcts.mcts_fixed_min(o, number, seed=seed, eval_id=e) == eg["mc_min"][e].tolist()
        assert mcts.mcts_fixed_moves(o, number, seed=seed, eval_id=e) == eg["mc_med"][e].tolist()
    o = int(eg["mc_origins"][0])
    assert mcts.mcts_fixed_moves(o, 7, seed=4243, eval_id=0) == eg["mc7_med"].tolist()
    assert mcts.mcts_fixed_min(o, 7, seed=4243, eval_id=0) == eg["mc7_min"].tolist()
    assert mcts.mcts_fixed_batch is mcts.mcts_fixed


def test_play_mcts_fixed_matches_oracle_main_loop(mcts, orc):
    # config 1: one full main game, 10 playouts per legal move, against the same loop on the oracle
    seed = 2718
    got = mcts.play_mcts_fixed(number=4, mode=2, seed=seed)
    board = orc.init_board(orc.rng_key(seed, 0))
    score = count = 0
    while True:
        res = orc.mcts_fixed_min(board, 4, seed, count + 1)
        i = int(np.argmax(res))
        f, s, m = orc.move(board, i)
        if not m:
            break
        board = orc.spawn(f, orc.rng_key(seed, 0), count)
        score += s
        count += 1
    assert got == (board, score, count)
    assert count > 50


def test_global_random_seed_reproducibility(board):
    random.seed(5)
    a = board.play_fixed_batch(64)
    random.seed(5)
    b = board.play_fixed_batch(64)
    assert a == b and len(a) == 64


def test_bare_name_shims_run_reference_style_script(tmp_path):
    # the reference's scripts import by bare module name (selfplay.py:3-6)
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from board import *\n"
        "from mcts import mcts_fixed_min\n"
        "from network import ConvNet\n"
        "random.seed(7)\n"
        "b = generate_init_tiles()\n"
        "r = mcts_fixed_min(b, 4)\n"
        "f, s, m = move(b, int(np.argmax(r)))\n"
        "assert m and len(r) == 4\n"
        "arr = BoardArray([generate_tile(f) for _ in range(5)])\n"
        "while arr.boards:\n"
        "    arr.move_batch([(0, 1, 3, 2)] * len(arr.boards))\n"
        "print('ok', sum(p.numel() for p in ConvNet(64, 3).parameters()))\n" % os.path.join(root, "2048nn_b200", "dropin"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(tmp_path), timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "ok 231820" in out.stdout
