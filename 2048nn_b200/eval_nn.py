"""Drop-in for the reference's `eval_nn` module (fqjin/2048NN eval_nn.py): policy games played
to the end (eval_nn) or to the first death (eval_nn_min), one persistent-kernel launch each."""
import numpy as np
import torch

from . import engine
from .board import *  # noqa: F401,F403  (eval_nn.py:1)
from .board import _host, _raise_if_flag, _session_seed
from .network import blob_for


def eval_nn(model, name=None, origin=None, number=100, device='cpu', verbose=False, seed=None, precision=None):
    """eval_nn.py:4-58: `number` policy games (fresh boards, or all from `origin`).
    Returns (mean log10(score+1), max score, mean moves), or writes models/{name}.npz with
    scores, moves, boards in death order like the reference."""
    blob, prec = blob_for(model, precision)
    start = None
    if origin is not None:
        if origin == 0:
            raise ValueError("empty range for randrange()")
        start = torch.full((number,), int(np.array(origin, dtype=np.uint64).view(np.int64)), dtype=torch.int64, device="cuda")
    r = engine.playout_nn(blob, prec, int(number), _session_seed(seed), 0, start=start)
    _raise_if_flag()
    moves = r["moves"].cpu().numpy().astype(np.int64)
    order = np.argsort(moves, kind="stable")  # death order: by step, then list position
    scores = r["score"].cpu().numpy().view(np.uint32).astype(np.int64)[order]
    dead_boards = _host(r["final"])[order]
    moves = moves[order]
    if name is None:
        return np.mean(np.log10(scores + 1)), np.amax(scores), np.mean(moves)
    np.savez('models/{}.npz'.format(name), scores=scores, moves=moves, boards=dead_boards)
    print(name)
    print(np.mean(np.log10(scores + 1)), np.amax(scores), np.mean(moves))


def eval_nn_min(model, number=50, repeats=4, device='cpu', seed=None, precision=None):
    """eval_nn.py:61-95: mean over `repeats` batches of the (0-based) step of the first death."""
    blob, prec = blob_for(model, precision)
    r = engine.playout_nn(blob, prec, int(number) * int(repeats), _session_seed(seed), 0, group_size=int(number),
                          want_final=False)
    _raise_if_flag()
    return np.mean(r["group_min"].cpu().numpy())
