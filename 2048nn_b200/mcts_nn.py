"""Drop-in for the reference's `mcts_nn` module (fqjin/2048NN mcts_nn.py): NN-policy root
search and greedy policy games, each call one persistent-kernel launch on the B200.

`model` may be this package's `network.ConvNet` or any torch module carrying the reference's
c64b3 state_dict.  `device` is accepted for signature compatibility; the computation always
runs on the current CUDA device (there is no CPU path)."""
import numpy as np

from . import board as _board
from . import engine
from .board import *  # noqa: F401,F403  (mcts_nn.py:1)
from .board import _dev, _host, _raise_if_flag, _session_seed
from .network import blob_for


def mcts_nn(model, origin, number=10, device='cpu', seed=None, eval_id=0, precision=None):
    """mcts_nn.py:4-43: per root move [Left, Up, Right, Down], 1 + the fewest policy moves any
    of `number` lines survives (all lines stop at the first death); 0 for an illegal move."""
    blob, prec = blob_for(model, precision)
    r = engine.mcts_nn_lines(blob, prec, _dev([origin]), int(number), _session_seed(seed), eval_id)
    _raise_if_flag()
    legal = r["legal"].cpu().numpy()[0]
    mn = r["minmov"].cpu().numpy()[0]
    return [int(mn[i]) + 1 if legal[i] else 0 for i in range(4)]


def play_nn(model, game=None, press_enter=False, device='cpu', verbose=False, seed=None, gid=0, precision=None):
    """mcts_nn.py:46-94: one greedy policy game -> (board, score, count)."""
    blob, prec = blob_for(model, precision)
    s = _session_seed(seed)
    if press_enter:
        return _play_nn_stepwise(blob, prec, game, s, gid)
    start = None if game is None else _dev([game])  # mcts_nn.py:64 uses `is None`, board 0 is a board
    if game == 0:
        raise ValueError("empty range for randrange()")  # the reference fails on an empty board's first spawn
    r = engine.playout_nn(blob, prec, 1, s, gid, start=start)
    _raise_if_flag()
    return int(_host(r["final"])[0]), int(r["score"].item()) & 0xFFFFFFFF, int(r["moves"].item())


def _play_nn_stepwise(blob, prec, game, seed, gid):
    """mcts_nn.py:68-91, the console mode of play_nn: one policy move per key press ('q' quits, and the reference then
    falls off the end of the function: None), the log-probs, the arrow and the board printed after every move.  Same
    spawn stream as the one-launch path, so a game stepped through to the end is the game play_nn returns."""
    import torch

    from . import _lib as L
    from .network import PRECISIONS

    if game is None:
        game = generate_init_tiles(seed, gid)  # noqa: F405
    arr = BoardArray([game], seed=seed, gid_base=gid)  # noqa: F405
    draw(game, 0)  # noqa: F405
    count = 0
    while True:
        if input() == 'q':
            return None
        b = _dev(arr.boards)
        out = torch.empty((1, 4), dtype=torch.float32, device=b.device)
        L.check(L.load_library().b2048_convnet_forward(L.ctx(b.device), PRECISIONS[prec], L.ptr(blob), L.ptr(b), L.ptr(out), 1,
                                                       L.stream_ptr(b.device)))
        order = torch.argsort(out, dim=1, descending=True, stable=True).cpu()  # mcts_nn.py:79
        before = arr.boards[0]
        dead = arr.move_batch(order)
        if dead:
            return before, dead[0], count
        count += 1
        played = next(int(i) for i in order[0] if _board.move(before, int(i))[2])
        print(out.cpu())
        print(ARROWS[played])  # noqa: F405
        draw(arr.boards[0], arr.scores[0])  # noqa: F405
