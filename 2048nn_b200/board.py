"""Drop-in for the reference's `board` module (fqjin/2048NN board.py), running on the B200.

Same names, argument meaning and return types as the reference: boards are Python ints,
`move` returns `(board, score, moved)`, `play_fixed` returns `(board, score, count)`,
`BoardArray.move_batch` replaces `self.boards/self.scores` with new lists, etc.  Every
computation is a CUDA kernel behind the C ABI (include/b2048.h); there is no CPU path.

Randomness.  The reference draws spawns from CPython's global `random` (board.py:6,70-98).
Here each top-level call draws ONE 64-bit session seed from that same global `random`
(`random.getrandbits(64)`), so `random.seed(k)` (board.py:8-12, selfplay.py:126-131) still makes
a run reproducible; the spawns themselves come from the counter-based per-game streams of
include/b2048.h.  Pass `seed=` / `gid_base=` explicitly for stream-exact replays (tests).
"""
import os
import random

import numpy as np
import torch

from . import _lib as L
from . import engine

# Seed (board.py:8-12).  The bare-name shim dropin/board.py also reseeds the global RNGs.
seed = 12345

# Print options / console helpers (board.py:14-20, 36-43): thin passthroughs
CLEAR = 'clear' if os.name == 'posix' else 'cls'
ARROWS = {0: '  ⮜                   .',
          1: '       ⮝              .',
          2: '            ⮞         .',
          3: '                 ⮟    .'}

LUDR = (0, 1, 3, 2)
_M64 = (1 << 64) - 1


def _session_seed(explicit=None):
    return (int(explicit) & _M64) if explicit is not None else random.getrandbits(64)


def _dev(x):
    """Python int(s) -> int64 device tensor"""
    return L.boards_to_device(x)


def _host(t):
    return L.boards_to_host(t)


def _raise_if_flag():
    if L.take_error_flag():
        raise ValueError("empty range for randrange()")  # what board.py:87 raises via randrange(0)


def get_tiles(x):
    """board.py:23-33: list of the 16 tile exponents, nibble 0 first."""
    return [(int(x) >> (4 * k)) & 0xF for k in range(16)]


def draw(x, score=None):
    """board.py:36-43 console rendering (debug passthrough)."""
    expo = np.power(2, get_tiles(x)).reshape(4, 4)
    expo = str(expo)
    print(expo.replace('1 ', '. ', 12).replace('1]', '.]', 4))
    print(' Score : {}'.format(score))


def transpose(x):
    """board.py:46-54"""
    t, _, _ = engine.board_ops(_dev([x]), True, False, False)
    return int(_host(t)[0])


def countzero(x):
    """board.py:57-65 (16 empties wrap to 0)"""
    _, cz, _ = engine.board_ops(_dev([x]), False, True, False)
    return int(cz.item())


def generate_init_tiles(seed=None, gid=0):
    """board.py:68-82: a fresh board with two tiles."""
    return int(_host(engine.init_boards(1, _session_seed(seed), gid))[0])


def generate_tile(board, seed=None, gid=0, k=0):
    """board.py:85-101: place a 2 (90%) or 4 (10%) on a uniformly random empty cell."""
    out, err = engine.spawn_batch(_dev([board]), _session_seed(seed), gid_base=gid, k=k)
    if int(err.item()):
        raise ValueError("empty range for randrange()")
    return int(_host(out)[0])


class _RowTable:
    """Lazy view of merge_table / merge_table_rev (board.py:138-148): table[row] -> (final, score, moved)."""

    def __init__(self, rev):
        self._rev, self._t = rev, None

    def _load(self):
        if self._t is None:
            t = engine.tables()
            fin = t["fin_rev" if self._rev else "fin_fwd"].cpu().numpy().view(np.uint16)
            sc = t["score"].cpu().numpy().view(np.uint32)
            mv = t["moved_rev" if self._rev else "moved_fwd"].cpu().numpy()
            self._t = (fin, sc, mv)
        return self._t

    def __len__(self):
        return 65536

    def __getitem__(self, row):
        fin, sc, mv = self._load()
        return int(fin[row]), int(sc[row]), bool(mv[row])


merge_table = _RowTable(False)
merge_table_rev = _RowTable(True)


def move_row(x, rev):
    """board.py:104-132 through the device-built tables."""
    return (merge_table_rev if rev else merge_table)[int(x) & 0xFFFF]


def move(x, direction):
    """board.py:151-189: (board, score, moved) for direction 0 L, 1 U, 2 R, 3 D (mod 4)."""
    ob, os_, om = engine.move_batch(_dev([x]), int(direction) & 3)
    return int(_host(ob)[0]), int(os_.item()), bool(om.item())


def play_fixed(board=None, press_enter=False, seed=None, gid=0):
    """board.py:192-225: one L,U,D,R game to the end -> (board, score, count)."""
    s = _session_seed(seed)
    if not press_enter:
        start = None if not board else _dev([board])
        f, sc, c = engine.playout_fixed(1, s, gid, start=start)
        _raise_if_flag()
        return int(_host(f)[0]), int(sc.item()) & 0xFFFFFFFF, int(c.item())
    # interactive debug path (board.py:209-222): one step per key press
    if not board:
        board = generate_init_tiles(s, gid)
    arr = BoardArray([board], seed=s, gid_base=gid)
    draw(board, 0)
    count = 0
    while True:
        if input() == 'q':
            break
        before = arr.boards[0]
        dead = arr.move_batch([LUDR])
        if dead:
            return before, dead[0], count
        count += 1
        draw(arr.boards[0], arr.scores[0])
    return arr.boards[0], arr.scores[0], count


class BoardArray:
    """board.py:228-303: a batch of games as parallel lists of boards and scores.

    Attributes:
        boards: list of boards (Python ints)
        scores: list of scores (start at 0: board.py:244)
    """

    def __init__(self, boards, seed=None, gid_base=0):
        self.boards = boards
        self.scores = [0] * len(boards)
        self._seed = _session_seed(seed)
        self._gids = np.arange(gid_base, gid_base + len(boards), dtype=np.uint64)
        self._k = np.zeros(len(boards), np.uint32)

    def _step(self, move_list):
        n = len(self.boards)
        if isinstance(move_list, torch.Tensor):
            prio = move_list.reshape(n, 4).to(device="cuda", dtype=torch.uint8).contiguous()
        else:
            prio = torch.from_numpy(np.ascontiguousarray(np.asarray(move_list, dtype=np.int64).reshape(n, 4) & 3,
                                                         dtype=np.uint8)).cuda()
        b = _dev(self.boards)
        gids = torch.from_numpy(self._gids.view(np.int64)).cuda()
        k = torch.from_numpy(self._k.view(np.int32)).cuda()
        ob, gain, moved = engine.policy_step(b, prio, self._seed, k, gids=gids)
        moved = moved.cpu().numpy().astype(bool)
        _raise_if_flag()
        return _host(ob), gain.cpu().numpy().view(np.uint32), moved

    def _keep(self, ob, gain, moved):
        self.boards = [int(x) for x in ob[moved]]
        self.scores = [int(s) + int(g) for s, g, m in zip(self.scores, gain, moved) if m]
        self._gids = self._gids[moved]
        self._k = self._k[moved] + np.uint32(1)

    def move_batch(self, move_list, get_boards=False):
        """board.py:246-277: one step; survivors stay in order, dead games are returned."""
        if not self.boards:
            return ([], []) if get_boards else []
        ob, gain, moved = self._step(move_list)
        dead_b = [int(b) for b, m in zip(self.boards, moved) if not m]
        dead_s = [int(s) for s, m in zip(self.scores, moved) if not m]
        self._keep(ob, gain, moved)
        if get_boards:
            return dead_s, dead_b
        return dead_s

    def fast_move_batch(self, move_list):
        """board.py:279-303: True as soon as any game cannot move (array left untouched)."""
        if not self.boards:
            return False
        ob, gain, moved = self._step(move_list)
        if not moved.all():
            return True
        self._keep(ob, gain, moved)
        return False


def play_fixed_batch(number, seed=None, gid_base=0):
    """board.py:306-315: `number` fresh L,U,D,R games -> list of scores in death order
    (by step, then by list position)."""
    if number <= 0:
        return []
    f, sc, c = engine.playout_fixed(int(number), _session_seed(seed), gid_base)
    _raise_if_flag()
    sc = sc.cpu().numpy().view(np.uint32)
    order = np.argsort(c.cpu().numpy(), kind="stable")
    return [int(x) for x in sc[order]]
