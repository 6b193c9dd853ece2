"""Drop-in for the reference's `mcts` module (fqjin/2048NN mcts.py): fixed-order root search.

`mcts_fixed` (the README's `mcts_fixed_batch`), `mcts_fixed_moves`, `mcts_fixed_min` and
`play_mcts_fixed` keep the reference's signatures and return conventions.  All 4 x `number`
playouts of a root evaluation run to the end inside ONE persistent kernel launch
(b2048_mcts_fixed); the reductions are applied to the per-line results with the reference's
own NumPy formulas so the values are bit-identical given identical spawns.
"""
import numpy as np

from . import board as _board
from . import engine
from .board import *  # noqa: F401,F403  (the reference does `from board import *`, mcts.py:1)
from .board import _dev, _raise_if_flag, _session_seed


def _lines(origin, number, seed, eval_id):
    r = engine.mcts_fixed_lines(_dev([origin]), int(number), _session_seed(seed), eval_id, fused=False)
    _raise_if_flag()
    legal = r["legal"].cpu().numpy()[0].astype(bool)
    rscore = r["rscore"].cpu().numpy()[0].view(np.uint32).astype(np.int64)
    lscore = r["lscore"].cpu().numpy()[0].view(np.uint32).astype(np.int64)
    lmoves = r["lmoves"].cpu().numpy()[0].astype(np.int64)
    return legal, rscore, lscore, lmoves


def mcts_fixed(origin, number=10, seed=None, eval_id=0):
    """mcts.py:4-34: mean log10(playout score + root score + 1) per root move
    [Left, Up, Right, Down]; -1 for an illegal move."""
    legal, rscore, lscore, lmoves = _lines(origin, number, seed, eval_id)
    result = []
    for i in range(4):
        if legal[i]:
            order = np.argsort(lmoves[i], kind="stable")  # the reference collects scores in death order
            scores = lscore[i][order]
            result.append(np.mean(np.log10(scores + rscore[i] + 1)))  # mcts.py:30-31
        else:
            result.append(-1)
    return result


mcts_fixed_batch = mcts_fixed  # README.md:11 names it so; the code name is mcts_fixed (mcts.py:4)


def mcts_fixed_moves(origin, number=10, seed=None, eval_id=0):
    """mcts.py:37-70: 1-based step at which cumulative deaths reach number//2; -1 if illegal."""
    legal, _, _, lmoves = _lines(origin, number, seed, eval_id)
    result = []
    for i in range(4):
        if legal[i]:
            died_at = np.sort(lmoves[i]) + 1
            h = number // 2
            result.append(int(died_at[h - 1 if h else 0]))
        else:
            result.append(-1)
    return result


def mcts_fixed_min(origin, number=10, seed=None, eval_id=0):
    """mcts.py:73-103: 1-based step of the first death among the lines; 0 if illegal."""
    legal, _, _, lmoves = _lines(origin, number, seed, eval_id)
    return [int(lmoves[i].min()) + 1 if legal[i] else 0 for i in range(4)]


def play_mcts_fixed(board=None, number=10, press_enter=False, mode=0, seed=None):
    """mcts.py:106-148: play a game choosing argmax of mcts_fixed / _moves / _min each move.

    The whole main game runs inside one kernel launch (b2048_selfplay_fixed: one CTA, no host round trip per main
    move) on the same spawn streams as the step-by-step loop below, which remains the interactive
    (`press_enter`) path.  Mode 0 compares fixed-point (2^-40) log-score means on the device where the loop uses
    NumPy's float64 mean: the two can only disagree when two root values differ by less than ~1e-11."""
    s = _session_seed(seed)
    if press_enter:
        return _play_mcts_fixed_host(board, number, True, mode, s)
    if mode not in (0, 1, 2):
        raise KeyError(mode)
    start = None if not board else _dev([board])
    max_steps = 8192
    while True:
        r = engine.selfplay_fixed_games(1, int(number), s, times=1, mode=int(mode), starts=start, max_steps=max_steps,
                                        record=False)
        _raise_if_flag()
        if int(r["status"].item()) == 0:
            break
        max_steps *= 8  # the game outlived the step budget: replay it with a larger one
    return int(_board._host(r["final"])[0]), int(r["score"].item()), int(r["steps"].item())


def _play_mcts_fixed_host(board=None, number=10, press_enter=False, mode=0, seed=None):
    """The reference's loop shape: one root evaluation (kernel launch) per main move, decision on the host."""
    s = _session_seed(seed)
    if not board:
        board = _board.generate_init_tiles(s, 0)
    score = 0
    count = 0
    if press_enter:
        draw(board, score)  # noqa: F405
    mcts_fn = {0: mcts_fixed, 1: mcts_fixed_moves, 2: mcts_fixed_min}[mode]
    while True:
        if press_enter and input() == 'q':
            break
        result = mcts_fn(board, number, seed=s, eval_id=count + 1)
        i = int(np.argmax(result))
        f, sc, m = _board.move(board, i)
        if not m:  # first argmax move is legal if any move is legal (mcts.py:138)
            break
        board = _board.generate_tile(f, seed=s, gid=0, k=count)
        score += sc
        count += 1
        if press_enter:
            print(ARROWS[i])  # noqa: F405
            print(result)
            draw(board, score)  # noqa: F405
    return board, score, count
