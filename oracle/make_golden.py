"""Generate tests/golden/*.npz by importing and running the UNMODIFIED Python reference.

Run in the build container only (needs /root/reference, read-only):

    python oracle/make_golden.py

The reference is imported from a scratch working directory (board.py:138-148 writes
./merge_tables.pickle into the CWD).  Spawn randomness is injected by rebinding the name
`randrange` in the reference's `board` module (looked up in module globals at call time,
board.py:6,70-98) to a per-game counter stream -- SURVEY.md A.3.  Every game is run one at a
time so the stream is per game.  The stream (mix64/splitmix) is restated here in pure Python,
independent of both the C oracle and the CUDA source.

Outputs (all small, committed):
  tests/golden/engine_golden.npz   row tables, move/transposes/countzero vectors, spawns,
                                   injected play_fixed games, mcts_fixed/_min/_moves results
  tests/golden/c64b3_golden.npz    shipped + random-init c64b3 state dicts, boards, fp32
                                   log-probs from network.ConvNet, policy games, mcts_nn results
"""
import os
import sys
import tempfile

import numpy as np

REF = os.environ.get("B2048_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
M64 = (1 << 64) - 1
GOLDEN = 0x9E3779B97F4A7C15
INIT_DRAWS = 8


def mix64(z):
    z &= M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def rng_key(seed, gid):
    return mix64(mix64((seed + GOLDEN) & M64) ^ ((gid * 0xD1342543DE82EF95 + 1) & M64))


def rng_word(key, w):
    d = mix64((key + ((w >> 1) + 1) * GOLDEN) & M64)
    return (d & 0xFFFFFFFF) if (w & 1) else (d >> 32)


class Stream:
    """Replacement for board.randrange: bound n -> (word * n) >> 32 on a seekable word cursor."""

    def __init__(self):
        self.key = 0
        self.w = 0

    def start(self, key, word=0):
        self.key, self.w = key, word

    def __call__(self, n):
        if n <= 0:
            raise ValueError("empty range for randrange()")
        v = (rng_word(self.key, self.w) * n) >> 32
        self.w += 1
        return v


def line_gid(eval_id, i, j, number):
    return (eval_id * 4 + i) * number + j


def main():
    scratch = tempfile.mkdtemp(prefix="b2048_ref_")
    os.chdir(scratch)
    sys.path.insert(0, REF)
    import board as rb  # the reference, unmodified
    import torch
    from network import ConvNet

    stream = Stream()
    rb.randrange = stream  # injected spawns
    rs = np.random.RandomState(20261019)
    os.makedirs(OUT, exist_ok=True)
    g = {}

    # ---- row tables (board.py:138-148) -------------------------------------------------
    g["fin_fwd"] = np.array([t[0] for t in rb.merge_table], np.uint16)
    g["fin_rev"] = np.array([t[0] for t in rb.merge_table_rev], np.uint16)
    g["score_fwd"] = np.array([t[1] for t in rb.merge_table], np.uint32)
    g["score_rev"] = np.array([t[1] for t in rb.merge_table_rev], np.uint32)
    g["moved_fwd"] = np.array([t[2] for t in rb.merge_table], np.uint8)
    g["moved_rev"] = np.array([t[2] for t in rb.merge_table_rev], np.uint8)

    # ---- move / transpose / countzero vectors (board.py:46-65,151-189) -----------------
    special = [0x2141348, 0x0123456789ABCDEF, 0xFFFFFFFFFFFFFFFF, 0x1111222233334444, 0, 1, 0xF, 0xFF,
               0xFF00000000000000, 0xF00F00000000F00F, 0xEEEEEEEEEEEEEEEE, 0xFEFEFEFEFEFEFEFE,
               0x1000000000000000, 0x0000000000010000]
    uni = rs.randint(0, 1 << 63, size=6000, dtype=np.int64).astype(np.uint64) * np.uint64(2) + \
        rs.randint(0, 2, size=6000).astype(np.uint64)
    # realistic: mostly small exponents, some empties
    nib = rs.choice(16, size=(8000, 16), p=np.array([6, 5, 5, 4, 3, 3, 2, 2, 1, 1, 1, 1, .5, .5, .5, .5]) / 36.0)
    real = np.zeros(8000, np.uint64)
    for k in range(16):
        real |= nib[:, k].astype(np.uint64) << np.uint64(4 * k)
    # high-tile boards (14/15 heavy) for the annihilation path
    nibh = rs.choice([0, 13, 14, 15], size=(2000, 16))
    high = np.zeros(2000, np.uint64)
    for k in range(16):
        high |= nibh[:, k].astype(np.uint64) << np.uint64(4 * k)
    boards = np.concatenate([np.array(special, np.uint64), uni, real, high])
    n = boards.size
    mv_b = np.zeros((4, n), np.uint64)
    mv_s = np.zeros((4, n), np.uint32)
    mv_m = np.zeros((4, n), np.uint8)
    tr = np.zeros(n, np.uint64)
    cz = np.zeros(n, np.uint8)
    for j, b in enumerate(boards.tolist()):
        for d in range(4):
            f, s, m = rb.move(b, d)
            mv_b[d, j], mv_s[d, j], mv_m[d, j] = f, s, m
        tr[j] = rb.transpose(b)
        cz[j] = rb.countzero(b)
    g["mv_boards"], g["mv_out"], g["mv_score"], g["mv_moved"] = boards, mv_b, mv_s, mv_m
    g["transpose"], g["countzero"] = tr, cz

    # ---- generate_tile with injected words (board.py:85-101) ---------------------------
    sp_in = [b for b in boards.tolist() if 0 < rb.countzero(b)][:6000]
    sp_key = rs.randint(0, 1 << 62, size=len(sp_in), dtype=np.int64).astype(np.uint64)
    sp_out = np.zeros(len(sp_in), np.uint64)
    for j, (b, k) in enumerate(zip(sp_in, sp_key.tolist())):
        stream.start(k, 2 * INIT_DRAWS)  # spawn 0 of stream k
        sp_out[j] = rb.generate_tile(b)
    g["sp_in"], g["sp_key"], g["sp_out"] = np.array(sp_in, np.uint64), sp_key, sp_out

    # ---- injected play_fixed games (board.py:68-82,192-225) ----------------------------
    def ref_fresh_game(seed, gid):
        key = rng_key(seed, gid)
        stream.start(key, 0)
        start = rb.generate_init_tiles()
        stream.start(key, 2 * INIT_DRAWS)
        f, s, c = rb.play_fixed(start)
        return start, f, s, c

    pf = np.array([ref_fresh_game(12345, gid) for gid in range(1500)], dtype=np.uint64)
    g["pf_seed"] = np.uint64(12345)
    g["pf_start"], g["pf_final"], g["pf_score"], g["pf_count"] = pf[:, 0], pf[:, 1], pf[:, 2], pf[:, 3]

    def ref_origin_game(origin, seed, gid):
        stream.start(rng_key(seed, gid), 2 * INIT_DRAWS)
        return rb.play_fixed(origin)

    po = np.array([ref_origin_game(0x2141348, 7, gid) for gid in range(500)], dtype=np.uint64)
    g["po_origin"], g["po_seed"] = np.uint64(0x2141348), np.uint64(7)
    g["po_final"], g["po_score"], g["po_count"] = po[:, 0], po[:, 1], po[:, 2]

    # ---- root search (mcts.py:4-103), one line at a time through the reference's own
    # move / generate_tile / BoardArray.move_batch; reductions by the reference's formulas --
    def ref_root(origin, number, seed, eval_id):
        scores = np.zeros((4, number), np.int64)
        moves = np.zeros((4, number), np.int64)
        legal, rscore = [], []
        for i in range(4):
            b, s, m = rb.move(origin, i)
            legal.append(m)
            rscore.append(s)
            if not m:
                continue
            for j in range(number):
                stream.start(rng_key(seed, line_gid(eval_id, i, j, number)), 2 * INIT_DRAWS)
                arr = rb.BoardArray([rb.generate_tile(b)])
                steps = 0
                while arr.boards:
                    dead = arr.move_batch([(0, 1, 3, 2)] * len(arr.boards))
                    steps += 1
                    if dead:
                        scores[i, j] = dead[0]
                moves[i, j] = steps - 1
        mean, mn, med = [], [], []
        for i in range(4):
            if legal[i]:
                mean.append(np.mean(np.log10(scores[i] + rscore[i] + 1)))  # mcts.py:30-31
                mn.append(1 + moves[i].min())                               # mcts.py:92-100
                h = number // 2                                             # mcts.py:57-67
                died_at = np.sort(moves[i] + 1)
                cum = [(c, int((died_at <= c).sum())) for c in np.unique(died_at)]
                med.append(next(c for c, k in cum if k >= h))
            else:
                mean.append(-1)
                mn.append(0)
                med.append(-1)
        return mean, mn, med, scores, moves

    origins = [0x2141348, 0x0000000100000001, 0x1111222233334444, 0x0123456789ABCDEF, 0x21, 0x0000000000102100]
    origins += [int(x) for x in g["pf_final"][:4]]  # dead boards: all moves illegal
    # mid-game boards from fixed playouts
    for gid in range(6):
        key = rng_key(99, gid)
        stream.start(key, 0)
        b = rb.generate_init_tiles()
        stream.start(key, 2 * INIT_DRAWS)
        for _ in range(40 + 25 * gid):
            for d in (0, 1, 3, 2):
                f, s, m = rb.move(b, d)
                if m:
                    b = rb.generate_tile(f)
                    break
        origins.append(b)
    number = 10
    mc_mean = np.zeros((len(origins), 4), np.float64)
    mc_min = np.zeros((len(origins), 4), np.int32)
    mc_med = np.zeros((len(origins), 4), np.int32)
    mc_sc = np.zeros((len(origins), 4, number), np.int64)
    mc_mv = np.zeros((len(origins), 4, number), np.int64)
    for e, o in enumerate(origins):
        mean, mn, med, sc, mv = ref_root(o, number, 4242, e)
        mc_mean[e], mc_min[e], mc_med[e], mc_sc[e], mc_mv[e] = mean, mn, med, sc, mv
    g["mc_origins"], g["mc_number"], g["mc_seed"] = np.array(origins, np.uint64), np.int64(number), np.uint64(4242)
    g["mc_mean"], g["mc_min"], g["mc_med"], g["mc_scores"], g["mc_moves"] = mc_mean, mc_min, mc_med, mc_sc, mc_mv
    # odd `number` exercises number//2 rounding of mcts_fixed_moves
    mean, mn, med, sc, mv = ref_root(origins[0], 7, 4243, 0)
    g["mc7_mean"], g["mc7_min"], g["mc7_med"] = np.array(mean), np.array(mn, np.int32), np.array(med, np.int32)

    np.savez_compressed(os.path.join(OUT, "engine_golden.npz"), **g)
    print("engine_golden.npz written", {k: getattr(v, "shape", None) for k, v in list(g.items())[:4]})

    # =====================================================================================
    # Network goldens (network.py:43-87, mcts_nn.py:4-43, eval_nn.py:4-58)
    # =====================================================================================
    n = {}
    ckpt = os.path.join(REF, "models/20200213/0_3400_soft3.0c64b3_p10_bs2048lr0.1d0.0_s0_best.pt")
    shipped = ConvNet(channels=64, blocks=3)
    shipped.load_state_dict(torch.load(ckpt, map_location="cpu"))
    shipped.eval()
    torch.manual_seed(0)
    rnd = ConvNet(channels=64, blocks=3)
    # fresh BN running stats are (0,1); make them non-trivial but deterministic
    with torch.no_grad():
        for mod in rnd.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.normal_(0, 0.5)
                mod.running_var.uniform_(0.5, 2.0)
                mod.weight.uniform_(0.5, 1.5)
                mod.bias.normal_(0, 0.2)
    rnd.eval()
    for tag, model in (("shipped", shipped), ("random", rnd)):
        for k, v in model.state_dict().items():
            n[f"sd_{tag}/{k}"] = v.numpy()

    def encode(bs):  # the reference's inline encoder, mcts_nn.py:25-35
        b = np.array(bs, dtype=np.uint64)
        data = []
        for _ in range(16):
            data.append(b & 0xF)
            b >>= 4
        t = torch.tensor(np.array(data), dtype=torch.float32).transpose(0, 1)
        t = [t == i for i in range(16)]
        return torch.stack(t, dim=1).float().view(-1, 16, 4, 4)

    # policy games with the shipped net, one game at a time, injected spawns
    def ref_policy_game(model, start, key, spawn0, record=None):
        stream.start(key, 2 * (INIT_DRAWS + spawn0))
        arr = rb.BoardArray([start])
        steps = 0
        final, score = start, 0
        with torch.no_grad():
            while arr.boards:
                if record is not None:
                    record.append(arr.boards[0])
                final = arr.boards[0]
                preds = torch.argsort(model(encode(arr.boards)), dim=1, descending=True)
                ds, db = arr.move_batch(preds, get_boards=True)
                steps += 1
                if ds:
                    score, final = ds[0], db[0]
        return final, score, steps - 1

    traj = []
    pg = []
    for gid in range(24):
        key = rng_key(2020, gid)
        stream.start(key, 0)
        start = rb.generate_init_tiles()
        f, s, c = ref_policy_game(shipped, start, key, 0, traj)
        pg.append((start, f, s, c))
        print("policy game", gid, s, c, flush=True)
    pg = np.array(pg, dtype=np.uint64)
    n["pg_seed"] = np.uint64(2020)
    n["pg_start"], n["pg_final"], n["pg_score"], n["pg_moves"] = pg[:, 0], pg[:, 1], pg[:, 2], pg[:, 3]

    # boards for logits parity: NN-policy trajectory boards + fixed-policy boards + uniform nibbles
    traj = np.array(traj, np.uint64)
    pick = rs.choice(traj.size, size=min(6000, traj.size), replace=False)
    fixed_tr = []
    for gid in range(12):
        key = rng_key(31, gid)
        stream.start(key, 0)
        b = rb.generate_init_tiles()
        stream.start(key, 2 * INIT_DRAWS)
        while True:
            fixed_tr.append(b)
            for d in (0, 1, 3, 2):
                f, s, m = rb.move(b, d)
                if m:
                    b = rb.generate_tile(f)
                    break
            else:
                break
    lb = np.concatenate([traj[pick], np.array(fixed_tr, np.uint64), uni[:512], np.array(special, np.uint64)])
    n["lg_boards"] = lb
    with torch.no_grad():
        n["lg_shipped"] = shipped(encode(lb)).numpy()
        n["lg_random"] = rnd(encode(lb)).numpy()

    # mcts_nn (mcts_nn.py:4-43): 1 + min_j L_j per legal root move, 0 if illegal
    def ref_mcts_nn(model, origin, number, seed, eval_id):
        out = []
        for i in range(4):
            b, _, m = rb.move(origin, i)
            if not m:
                out.append(0)
                continue
            best = None
            for j in range(number):
                key = rng_key(seed, line_gid(eval_id, i, j, number))
                stream.start(key, 2 * INIT_DRAWS)
                start = rb.generate_tile(b)
                _, _, L = ref_policy_game(model, start, key, 1)
                best = L if best is None else min(best, L)
            out.append(1 + best)
        return out

    nn_origins = [int(traj[k]) for k in (50, 400, 900, 1500)] + [int(pg[0, 1])]
    res = []
    for e, o in enumerate(nn_origins):
        res.append(ref_mcts_nn(rnd, o, 6, 555, e))  # random net: short lines, cheap
        print("mcts_nn", e, res[-1], flush=True)
    n["mn_origins"], n["mn_number"], n["mn_seed"] = np.array(nn_origins, np.uint64), np.int64(6), np.uint64(555)
    n["mn_result_random"] = np.array(res, np.int32)

    np.savez_compressed(os.path.join(OUT, "c64b3_golden.npz"), **n)
    print("c64b3_golden.npz written; logits boards:", lb.size)


if __name__ == "__main__":
    main()
