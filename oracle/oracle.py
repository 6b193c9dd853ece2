"""ctypes front-end of the CPU oracle (oracle/b2048_oracle.c) + the fp32 network oracle.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (2048nn_b200/) never imports it.

Parity status: PINNED against golden vectors generated from the unmodified Python reference
by oracle/make_golden.py (tests/test_oracle_golden.py).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libb2048_oracle.so")

u64 = C.c_uint64
u32 = C.c_uint32
P = C.c_void_p


def build(force=False):
    """Compile the C restatement (gcc, a second or two)."""
    src = os.path.join(_HERE, "b2048_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"], env={**os.environ, "ORC_CC": "gcc"})
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_transpose.restype = u64
        L.orc_transpose.argtypes = [u64]
        L.orc_countzero.restype = C.c_int
        L.orc_countzero.argtypes = [u64]
        L.orc_rng_key.restype = u64
        L.orc_rng_key.argtypes = [u64, u64]
        L.orc_rng_draw.restype = u64
        L.orc_rng_draw.argtypes = [u64, u64]
        L.orc_rng_word.restype = u32
        L.orc_rng_word.argtypes = [u64, u64]
        L.orc_generate_tile.restype = u64
        L.orc_generate_tile.argtypes = [u64, u32, u32, P]
        L.orc_init_board.restype = u64
        L.orc_init_board.argtypes = [u64]
        L.orc_spawn.restype = u64
        L.orc_spawn.argtypes = [u64, u64, u64, P]
        L.orc_move.argtypes = [u64, C.c_int, P, P, P]
        L.orc_move_batch.argtypes = [P, P, C.c_int, P, P, P, C.c_long]
        L.orc_play_order.argtypes = [u64, P, u64, u64, P, P, P]
        L.orc_play_fixed_batch.argtypes = [C.c_long, P, u64, u64, P, P, P, C.c_int]
        L.orc_line_gid.restype = u64
        L.orc_line_gid.argtypes = [u64, C.c_int, u64, u64]
        L.orc_root_lines.argtypes = [u64, u64, u64, u64, P, P, P, P, P, P]
        L.orc_mcts_fixed.argtypes = [u64, u64, u64, u64, P, P]
        L.orc_mcts_fixed_min.argtypes = [u64, u64, u64, u64, P]
        L.orc_mcts_fixed_moves.argtypes = [u64, u64, u64, u64, P]
        L.orc_policy_step.argtypes = [C.c_long, P, P, P, P, P, P]
        L.orc_legal_mask.argtypes = [u64]
        L.orc_init()
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(P) if a is not None else None


LUDR = (0, 1, 3, 2)


def tables():
    """(fin_fwd u16, fin_rev u16, score_fwd u32, score_rev u32, moved_fwd u8, moved_rev u8)"""
    out = [np.empty(65536, d) for d in (np.uint16, np.uint16, np.uint32, np.uint32, np.uint8, np.uint8)]
    lib().orc_tables(*[_p(a) for a in out])
    return out


def transpose(x):
    return int(lib().orc_transpose(int(x)))


def countzero(x):
    return int(lib().orc_countzero(int(x)))


def move(x, direction):
    b, s, m = u64(), u32(), C.c_uint8()
    lib().orc_move(int(x), int(direction), C.byref(b), C.byref(s), C.byref(m))
    return int(b.value), int(s.value), bool(m.value)


def move_batch(boards, dirs):
    """boards u64[N]; dirs: int or u8[N] -> (boards u64[N], score u32[N], moved u8[N])"""
    boards = np.ascontiguousarray(boards, np.uint64)
    n = boards.size
    ob, os_, om = np.empty(n, np.uint64), np.empty(n, np.uint32), np.empty(n, np.uint8)
    if np.isscalar(dirs):
        lib().orc_move_batch(_p(boards), None, int(dirs), _p(ob), _p(os_), _p(om), n)
    else:
        d = np.ascontiguousarray(dirs, np.uint8)
        lib().orc_move_batch(_p(boards), _p(d), 0, _p(ob), _p(os_), _p(om), n)
    return ob, os_, om


def rng_key(seed, gid):
    return int(lib().orc_rng_key(int(seed), int(gid)))


def rng_draw(key, i):
    return int(lib().orc_rng_draw(int(key), int(i)))


def rng_word(key, w):
    return int(lib().orc_rng_word(int(key), int(w)))


def generate_tile(board, w_pos, w_ten):
    err = C.c_int(0)
    r = lib().orc_generate_tile(int(board), int(w_pos), int(w_ten), C.byref(err))
    if err.value:
        raise ValueError("empty range for randrange()")  # what the reference raises
    return int(r)


def init_board(key):
    return int(lib().orc_init_board(int(key)))


def spawn(board, key, k):
    err = C.c_int(0)
    r = lib().orc_spawn(int(board), int(key), int(k), C.byref(err))
    if err.value:
        raise ValueError("empty range for randrange()")
    return int(r)


def play_order(start, order, key, spawn0=0):
    o = (C.c_uint8 * 4)(*order)
    f, s, c = u64(), u64(), u32()
    rc = lib().orc_play_order(int(start), o, int(key), int(spawn0), C.byref(f), C.byref(s), C.byref(c))
    if rc:
        raise ValueError("empty range for randrange()")
    return int(f.value), int(s.value), int(c.value)


def play_fixed_batch(n, seed, gid_base=0, starts=None, nthreads=1):
    """-> finals u64[n], scores u64[n], counts u32[n], indexed by game id"""
    f, s, c = np.empty(n, np.uint64), np.empty(n, np.uint64), np.empty(n, np.uint32)
    st = None if starts is None else np.ascontiguousarray(starts, np.uint64)
    rc = lib().orc_play_fixed_batch(n, _p(st), int(seed), int(gid_base), _p(f), _p(s), _p(c), int(nthreads))
    if rc:
        raise ValueError("empty range for randrange()")
    return f, s, c


def line_gid(eval_id, root_move, line, number):
    return int(lib().orc_line_gid(int(eval_id), int(root_move), int(line), int(number)))


def root_lines(origin, number, seed, eval_id, order=LUDR):
    o = (C.c_uint8 * 4)(*order)
    legal = np.zeros(4, np.uint8)
    rs = np.zeros(4, np.uint32)
    sc = np.zeros(4 * number, np.uint64)
    mv = np.zeros(4 * number, np.uint32)
    fin = np.zeros(4 * number, np.uint64)
    rc = lib().orc_root_lines(int(origin), number, int(seed), int(eval_id), o, _p(legal), _p(rs), _p(sc), _p(mv), _p(fin))
    if rc:
        raise ValueError("empty range for randrange()")
    return legal, rs, sc.reshape(4, number), mv.reshape(4, number), fin.reshape(4, number)


def mcts_fixed(origin, number=10, seed=0, eval_id=0):
    """mcts.py:4-34 with per-line streams; np.mean applied to the oracle's log10 terms."""
    legal, rs, sc, mv, _ = root_lines(origin, number, seed, eval_id)
    out = []
    for i in range(4):
        if legal[i]:
            out.append(np.mean(np.log10(sc[i].astype(np.int64) + int(rs[i]) + 1)))
        else:
            out.append(-1)
    return out


def mcts_fixed_min(origin, number=10, seed=0, eval_id=0):
    r = (C.c_int * 4)()
    lib().orc_mcts_fixed_min(int(origin), number, int(seed), int(eval_id), r)
    return list(r)


def mcts_fixed_moves(origin, number=10, seed=0, eval_id=0):
    r = (C.c_int * 4)()
    lib().orc_mcts_fixed_moves(int(origin), number, int(seed), int(eval_id), r)
    return list(r)


def legal_mask(board):
    return int(lib().orc_legal_mask(int(board)))


def policy_step(boards, scores, spawn_idx, keys, prio, alive):
    """In-place BoardArray.move_batch step with per-game priority rows (n x 4 u8)."""
    n = boards.size
    prio = np.ascontiguousarray(prio, np.uint8)
    return lib().orc_policy_step(n, _p(boards), _p(scores), _p(spawn_idx), _p(keys), _p(prio), _p(alive))


# --------------------------------------------------------------------------------------
# Network oracle: plain PyTorch fp32 restatement of network.py:43-87 (ConvNet c64b3) and of
# the inline one-hot encoder (mcts_nn.py:25-35).  Functional form over a state_dict with the
# reference's key names.  Floating point: this is the fp32 reference.
# --------------------------------------------------------------------------------------
def onehot(boards):
    """u64[N] -> float32 (N,16,4,4): channel c is 1 where nibble == c (mcts_nn.py:25-35)."""
    import torch

    b = np.ascontiguousarray(boards, np.uint64)
    nib = np.stack([(b >> np.uint64(4 * k)) & np.uint64(0xF) for k in range(16)], axis=1).astype(np.int64)
    t = torch.from_numpy(nib)  # (N,16) position-major
    oh = torch.stack([(t == c) for c in range(16)], dim=1).float()  # (N,16ch,16pos)
    return oh.view(-1, 16, 4, 4)


def convnet_forward(sd, x, blocks=3, return_prelogits=False):
    """network.py:78-87.  sd: dict name -> torch tensor (fp32).  Eval-mode BN (running stats)."""
    import torch
    import torch.nn.functional as F

    def bn(x, pfx):
        return F.batch_norm(x, sd[pfx + ".running_mean"], sd[pfx + ".running_var"], sd[pfx + ".weight"],
                            sd[pfx + ".bias"], False, 0.0, 1e-5)

    with torch.no_grad():
        x = F.relu(bn(F.conv2d(x, sd["in_block.0.weight"], padding=1), "in_block.1"))
        for b in range(blocks):
            h = F.relu(bn(F.conv2d(x, sd[f"blocks.{b}.0.weight"], padding=1), f"blocks.{b}.1"))
            h = bn(F.conv2d(h, sd[f"blocks.{b}.3.weight"], padding=1), f"blocks.{b}.4")
            x = F.relu(x + h)
        x = F.relu(bn(F.conv2d(x, sd["out_block.0.weight"]), "out_block.1"))
        x = x.reshape(-1, sd["policy.weight"].shape[1])
        z = F.linear(x, sd["policy.weight"], sd["policy.bias"])
        if return_prelogits:
            return z
        return F.log_softmax(z, dim=1)


def policy_order(logp):
    """argsort(descending) as mcts_nn.py:36 (ties: unspecified in torch; stable here)."""
    return np.argsort(-np.asarray(logp), axis=1, kind="stable").astype(np.uint8)


def nn_games(sd, starts, keys, spawn0=0, stop_at_first_death=False, max_steps=1 << 20):
    """Policy games (eval_nn.py:28-47 / mcts_nn.py:24-39) with per-game injected streams.

    Returns finals u64, scores u64, moves u32 (L_j = accepted moves), steps run.
    stop_at_first_death: mcts_nn / eval_nn_min semantics -- stop after the first step in which
    some game fails to move (games that already moved in that step keep their new state;
    callers only use min(moves))."""
    import torch

    boards = np.array(starts, dtype=np.uint64)
    n = boards.size
    scores = np.zeros(n, np.uint64)
    sidx = np.full(n, spawn0, np.uint32)
    keys = np.ascontiguousarray(keys, np.uint64)
    alive = np.ones(n, np.uint8)
    moves = np.zeros(n, np.uint32)
    steps = 0
    while alive.any() and steps < max_steps:
        idx = np.nonzero(alive)[0]
        logp = convnet_forward(sd, onehot(boards[idx])).numpy()
        prio = np.zeros((n, 4), np.uint8)
        prio[idx] = policy_order(logp)
        deaths = policy_step(boards, scores, sidx, keys, prio, alive)
        moves[alive.astype(bool)] += 1
        steps += 1
        if deaths and stop_at_first_death:
            break
    return boards, scores, moves, steps


def mcts_nn(sd, origin, number=10, seed=0, eval_id=0, want_moves=False):
    """mcts_nn.py:4-43 on per-line injected streams (line j of root move i of evaluation e is game
    (e*4+i)*number+j, first spawn k = 0): 1 + the fewest policy moves any line survives, 0 for an illegal move.
    want_moves: also return the number of accepted line moves played (the playout-moves metric)."""
    out, played = [], 0
    for i in range(4):
        b, _, m = move(origin, i)
        if not m:
            out.append(0)
            continue
        keys = np.array([rng_key(seed, line_gid(eval_id, i, j, number)) for j in range(number)], np.uint64)
        starts = np.array([spawn(b, int(k), 0) for k in keys], np.uint64)  # mcts_nn.py:22
        _, _, mv, _ = nn_games(sd, starts, keys, spawn0=1, stop_at_first_death=True)
        out.append(1 + int(mv.min()))
        played += int(mv.sum())
    return (out, played) if want_moves else out


def selfplay_nn(sd, G, number, seed, max_steps, game_id_base=0, games_total=None):
    """selfplay.py:62-115 for main games game_id_base..+G-1 of a session of games_total games, at most max_steps main
    moves each: list of (boards u64[T], results f32[T,4], score, final board)."""
    games_total = game_id_base + G if games_total is None else games_total
    recs = []
    for g in range(game_id_base, game_id_base + G):
        key = rng_key(seed, g)
        b, score, hb, hr = init_board(key), 0, [], []
        for step in range(max_steps):
            res = mcts_nn(sd, b, number, seed, (step + 1) * games_total + g)
            i = int(np.argmax(res))           # selfplay.py:95
            f, s, m = move(b, i)
            if not m:                          # selfplay.py:96-98
                break
            hb.append(b)
            hr.append(res)
            b = spawn(f, key, step)
            score += s
        recs.append((np.array(hb, np.uint64), np.array(hr, np.float32).reshape(-1, 4), score, b))
    return recs
