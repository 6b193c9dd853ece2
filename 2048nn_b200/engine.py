"""Tensor-level host API over the C ABI: device tensors in, device tensors out, no syncs.

Boards are torch.int64 tensors holding the uint64 bit pattern (nibble k = cell (k/4, k%4)).
Each function names the reference code it replaces (file:line in fqjin/2048NN).
"""
import ctypes as C
import os

import torch

from . import _lib as L

LUDR = (0, 1, 3, 2)  # board.py:214


def _dev(t):
    return t.device


def tables(device="cuda"):
    """Row tables built on the device (board.py:104-148): dict of device tensors."""
    d = torch.device(device)
    out = dict(fin_fwd=torch.empty(65536, dtype=torch.int16, device=d),
               fin_rev=torch.empty(65536, dtype=torch.int16, device=d),
               score=torch.empty(65536, dtype=torch.int32, device=d),
               moved_fwd=torch.empty(65536, dtype=torch.uint8, device=d),
               moved_rev=torch.empty(65536, dtype=torch.uint8, device=d))
    L.check(L.load_library().b2048_tables(L.ctx(d), L.ptr(out["fin_fwd"]), L.ptr(out["fin_rev"]), L.ptr(out["score"]),
                                          L.ptr(out["moved_fwd"]), L.ptr(out["moved_rev"]), L.stream_ptr(d)))
    return out


def move_batch(boards, dirs):
    """board.py:151-189 move() for N boards.  dirs: int or uint8 tensor[N].
    -> (boards int64[N], score int32[N], moved uint8[N])"""
    d = _dev(boards)
    n = boards.numel()
    ob = torch.empty_like(boards)
    os_ = torch.empty(n, dtype=torch.int32, device=d)
    om = torch.empty(n, dtype=torch.uint8, device=d)
    if isinstance(dirs, torch.Tensor):
        dirs = dirs.to(device=d, dtype=torch.uint8).contiguous()
        dp, ds = L.ptr(dirs), 0
    else:
        dp, ds = None, int(dirs)
    L.check(L.load_library().b2048_move_batch(L.ctx(d), L.ptr(boards), dp, ds, L.ptr(ob), L.ptr(os_), L.ptr(om), n,
                                              L.stream_ptr(d)))
    return ob, os_, om


def board_ops(boards, transpose=True, countzero=True, legal=True):
    """board.py:46-65 transpose / countzero, and the legal-move mask, for N boards."""
    d = _dev(boards)
    n = boards.numel()
    t = torch.empty_like(boards) if transpose else None
    cz = torch.empty(n, dtype=torch.uint8, device=d) if countzero else None
    lg = torch.empty(n, dtype=torch.uint8, device=d) if legal else None
    L.check(L.load_library().b2048_board_ops(L.ctx(d), L.ptr(boards), L.ptr(t), L.ptr(cz), L.ptr(lg), n, L.stream_ptr(d)))
    return t, cz, lg


def init_boards(n, seed, gid_base=0, device="cuda"):
    """board.py:68-82 generate_init_tiles for games gid_base..gid_base+n-1."""
    L.require_cuda()
    d = torch.device(device)
    out = torch.empty(n, dtype=torch.int64, device=d)
    L.check(L.load_library().b2048_init_boards(L.ctx(d), seed, gid_base, L.ptr(out), n, L.stream_ptr(d)))
    return out


def spawn_batch(boards, seed, gids=None, gid_base=0, k=0):
    """board.py:85-101 generate_tile: spawn number k of stream (seed, gid) on each board.
    Returns (boards, err) where err is an int32[1] device flag (countzero == 0 somewhere)."""
    d = _dev(boards)
    n = boards.numel()
    out = torch.empty_like(boards)
    err = torch.zeros(1, dtype=torch.int32, device=d)
    if isinstance(k, torch.Tensor):
        k = k.to(device=d, dtype=torch.int32).contiguous()
        kp, ks = L.ptr(k), 0
    else:
        kp, ks = None, int(k)
    L.check(L.load_library().b2048_spawn_batch(L.ctx(d), L.ptr(boards), seed, L.ptr(gids), gid_base, kp, ks, L.ptr(out),
                                               L.ptr(err), n, L.stream_ptr(d)))
    return out, err


def policy_step(boards, prio, seed, k, gids=None, gid_base=0):
    """board.py:246-303: one move_batch step with per-game priority rows prio uint8[N,4].
    -> (boards int64[N], gain int32[N], moved uint8[N])"""
    d = _dev(boards)
    n = boards.numel()
    ob = torch.empty_like(boards)
    gain = torch.empty(n, dtype=torch.int32, device=d)
    moved = torch.empty(n, dtype=torch.uint8, device=d)
    L.check(L.load_library().b2048_policy_step(L.ctx(d), L.ptr(boards), L.ptr(prio), seed, L.ptr(gids), gid_base, L.ptr(k),
                                               L.ptr(ob), L.ptr(gain), L.ptr(moved), n, L.stream_ptr(d)))
    return ob, gain, moved


def encode_boards(boards, onehot=True):
    """game_dataset.py:15-21 / 46-48: int64 boards -> float32 (N,16) exponents or (N,16,4,4) one-hot."""
    d = _dev(boards)
    n = boards.numel()
    out = torch.empty((n, 16, 4, 4) if onehot else (n, 16), dtype=torch.float32, device=d)
    L.check(L.load_library().b2048_encode_boards(L.ctx(d), L.ptr(boards), 1 if onehot else 0, L.ptr(out), n, L.stream_ptr(d)))
    return out


def soft_targets(results, soft=3.0):
    """game_dataset.py:25-28: softmax(results / rowmax * soft), float32 (N,4) device tensor."""
    d = _dev(results)
    results = results.contiguous()
    out = torch.empty_like(results)
    L.check(L.load_library().b2048_soft_targets(L.ctx(d), L.ptr(results), float(soft), L.ptr(out), results.shape[0], L.stream_ptr(d)))
    return out


def rng_draws(n, seed, gid_base, first, count, device="cuda"):
    d = torch.device(device)
    out = torch.empty((n, count), dtype=torch.int64, device=d)
    L.check(L.load_library().b2048_rng_draws(L.ctx(d), seed, gid_base, first, count, L.ptr(out), n, L.stream_ptr(d)))
    return out


def playout_fixed(n, seed, gid_base=0, start=None, order=LUDR, spawn0=0, device="cuda", out=None):
    """K2: n fixed-order games to the end in one persistent kernel (board.py:192-225, 306-315).
    -> (final int64[n], score int32[n], moves int32[n]) device tensors, indexed by game."""
    L.require_cuda()
    d = torch.device(device) if start is None else _dev(start)
    if out is None:
        out = (torch.empty(n, dtype=torch.int64, device=d), torch.empty(n, dtype=torch.int32, device=d),
               torch.empty(n, dtype=torch.int32, device=d))
    o = (C.c_uint8 * 4)(*[int(x) & 3 for x in order])
    L.check(L.load_library().b2048_playout_fixed(L.ctx(d), L.ptr(start), o, seed, gid_base, spawn0, L.ptr(out[0]),
                                                 L.ptr(out[1]), L.ptr(out[2]), n, L.stream_ptr(d)))
    return out


def mcts_fixed_lines(origins, number, seed, eval_id_base=0, line_begin=0, line_end=None, per_line=True, fused=True):
    """Root search lines for G origins (mcts.py:4-103).  Returns a dict of device tensors:
    legal u8[G,4], rscore i32[G,4], and (per_line) lscore/lmoves i32[G,4,number],
    (fused) logsum i64[G,4] (fixed point 2^-40), count i32[G,4], minmov i32[G,4]."""
    d = _dev(origins)
    G = origins.numel()
    line_end = number if line_end is None else line_end
    r = dict(legal=torch.empty((G, 4), dtype=torch.uint8, device=d), rscore=torch.empty((G, 4), dtype=torch.int32, device=d))
    if per_line:
        r["lscore"] = torch.zeros((G, 4, number), dtype=torch.int32, device=d)
        r["lmoves"] = torch.zeros((G, 4, number), dtype=torch.int32, device=d)
    if fused:
        r["logsum"] = torch.zeros((G, 4), dtype=torch.int64, device=d)
        r["count"] = torch.zeros((G, 4), dtype=torch.int32, device=d)
        r["minmov"] = torch.full((G, 4), 2**31 - 1, dtype=torch.int32, device=d)
    L.check(L.load_library().b2048_mcts_fixed(
        L.ctx(d), L.ptr(origins), G, number, line_begin, line_end, seed, eval_id_base, L.ptr(r["legal"]), L.ptr(r["rscore"]),
        L.ptr(r.get("lscore")), L.ptr(r.get("lmoves")), L.ptr(r.get("logsum")), L.ptr(r.get("count")), L.ptr(r.get("minmov")),
        L.stream_ptr(d)))
    return r


PRECISIONS = {"fp32": 0, "fp16": 1, "bf16": 2, "fp16x2": 3}


def playout_nn(blob, precision, n, seed, gid_base=0, start=None, spawn0=0, group_size=0, device="cuda", want_final=True):
    """K4: n NN-policy games in one persistent kernel (eval_nn.py:4-95, mcts_nn.py:46-94).
    -> dict(final i64[n], score i32[n], moves i32[n], group_min i32[groups] or None, total_moves i64[1])"""
    L.require_cuda()
    d = torch.device(device) if start is None else _dev(start)
    r = dict(final=torch.empty(n, dtype=torch.int64, device=d) if want_final else None,
             score=torch.empty(n, dtype=torch.int32, device=d), moves=torch.empty(n, dtype=torch.int32, device=d),
             total_moves=torch.zeros(1, dtype=torch.int64, device=d), group_min=None)
    if group_size > 0:
        r["group_min"] = torch.full(((n + group_size - 1) // group_size,), 2**31 - 1, dtype=torch.int32, device=d)
    L.check(L.load_library().b2048_playout_nn(L.ctx(d), PRECISIONS[precision], L.ptr(blob), L.ptr(start), seed, gid_base,
                                              spawn0, group_size, L.ptr(r["group_min"]), L.ptr(r["final"]), L.ptr(r["score"]),
                                              L.ptr(r["moves"]), L.ptr(r["total_moves"]), n, L.stream_ptr(d)))
    return r


def playout_nn_population(blobs, precision, number, seed, gid_base=0, gid_model_stride=0, start=None, group_size=0,
                          want_final=True):
    """K4 for a population (evolve.py:30-36): blobs is a (M, blob) device tensor of weight images (one row per network);
    every model plays `number` policy games, all models side by side in one launch per <= num_sms models.
    -> dict(final i64[M,number], score i32[M,number], moves i32[M,number], group_min i32[M,number/group_size] or None,
            total_moves i64[1]).  Row m equals playout_nn(blobs[m], ..., gid_base + m*gid_model_stride)."""
    L.require_cuda()
    if blobs.dim() != 2 or not blobs.is_contiguous():
        raise ValueError("blobs must be a contiguous (models, blob) tensor of weight images")
    d = _dev(blobs)
    M = blobs.shape[0]
    r = dict(final=torch.empty((M, number), dtype=torch.int64, device=d) if want_final else None,
             score=torch.empty((M, number), dtype=torch.int32, device=d),
             moves=torch.empty((M, number), dtype=torch.int32, device=d),
             total_moves=torch.zeros(1, dtype=torch.int64, device=d), group_min=None)
    if group_size > 0:
        if number % group_size:
            raise ValueError("number must be a multiple of group_size")
        r["group_min"] = torch.full((M, number // group_size), 2**31 - 1, dtype=torch.int32, device=d)
    L.check(L.load_library().b2048_playout_nn_population(
        L.ctx(d), PRECISIONS[precision], L.ptr(blobs), blobs.stride(0) * blobs.element_size(), M, L.ptr(start), seed, gid_base, gid_model_stride,
        group_size, L.ptr(r["group_min"]), L.ptr(r["final"]), L.ptr(r["score"]), L.ptr(r["moves"]), L.ptr(r["total_moves"]),
        number, L.stream_ptr(d)))
    return r


def mcts_nn_lines(blob, precision, origins, number, seed, eval_id_base=0, line_begin=0, line_end=None, per_line=False):
    """mcts_nn.py:4-43 for G origins: legal u8[G,4], rscore i32[G,4], minmov i32[G,4] (INT32_MAX
    where no line was played), lmoves i32[G,4,number] when per_line, total_moves i64[1]."""
    d = _dev(origins)
    G = origins.numel()
    line_end = number if line_end is None else line_end
    r = dict(legal=torch.empty((G, 4), dtype=torch.uint8, device=d), rscore=torch.empty((G, 4), dtype=torch.int32, device=d),
             minmov=torch.full((G, 4), 2**31 - 1, dtype=torch.int32, device=d),
             total_moves=torch.zeros(1, dtype=torch.int64, device=d))
    if per_line:
        r["lmoves"] = torch.zeros((G, 4, number), dtype=torch.int32, device=d)
    L.check(L.load_library().b2048_mcts_nn(L.ctx(d), PRECISIONS[precision], L.ptr(blob), L.ptr(origins), G, number, line_begin,
                                           line_end, seed, eval_id_base, L.ptr(r["legal"]), L.ptr(r["rscore"]),
                                           L.ptr(r.get("lmoves")), L.ptr(r["minmov"]), L.ptr(r["total_moves"]),
                                           L.stream_ptr(d)))
    return r


class HostPlayout:
    """Host-buffer front end of K2 for large batches: pinned host start boards in, pinned host
    (score, moves[, final]) out.  The batch is cut into `chunks` pieces that flow through three streams
    (H2D copy, kernel, D2H copy) so PCIe traffic hides behind the playouts.  Buffers are allocated once
    and reused (what a caller streaming batches would do)."""

    def __init__(self, n, device="cuda", chunks=5, want_final=False, ramp=True):
        L.require_cuda()
        self.n = n
        self.device = torch.device(device)
        self.want_final = want_final
        self.chunks = max(1, min(chunks, n))
        # ramp: small first and last pieces (triangular sizes 1:3:5:3:1), so that the first copy in and the last copy out --
        # the only transfers that nothing hides -- are short.  Measured at 16 M games (profiles/r02_e2e_chunks.txt): 5 ramped
        # pieces 20.9 ms per step, 4-6 equal pieces 21.6 ms, device-resident 19.6 ms; every further piece costs ~0.2 ms.
        w = [min(i + 1, self.chunks - i) * 2 - 1 for i in range(self.chunks)] if ramp else [1] * self.chunks
        cum = [0]
        for x in w:
            cum.append(cum[-1] + x)
        self.bounds = [(cum[i] * n // cum[-1], cum[i + 1] * n // cum[-1]) for i in range(self.chunks)]
        self.bounds = [(b, e) for b, e in self.bounds if e > b]
        self.d_start = torch.empty(n, dtype=torch.int64, device=self.device)
        self.d_out = (torch.empty(n, dtype=torch.int64, device=self.device) if want_final else None,
                      torch.empty(n, dtype=torch.int32, device=self.device),
                      torch.empty(n, dtype=torch.int32, device=self.device))
        self.h_out = (torch.empty(n, dtype=torch.int64).pin_memory() if want_final else None,
                      torch.empty(n, dtype=torch.int32).pin_memory(), torch.empty(n, dtype=torch.int32).pin_memory())
        # two alternating kernel streams: the tail of chunk i (a few long games) overlaps the head of chunk i+1
        self.s_in, self.s_out = torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)
        self.s_runs = [torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)]
        self.h2d_bytes = n * 8
        self.d2h_bytes = n * (16 if want_final else 8)

    def __call__(self, h_start, seed, gid_base=0, order=LUDR):
        """h_start: pinned int64[n] host tensor of start boards (0 = fresh board).  Returns the pinned host
        tensors (final or None, score, moves); the caller's stream waits for the result and the host is
        synchronised before returning."""
        cur = torch.cuda.current_stream(self.device)
        for st in (self.s_in, self.s_out, *self.s_runs):
            st.wait_stream(cur)
        o = (C.c_uint8 * 4)(*[int(x) & 3 for x in order])
        lib, ctx = L.load_library(), L.ctx(self.device)
        trace = [] if os.environ.get("B2048_E2E_TRACE") else None  # tools/e2e_diag.py: per-chunk event times
        for i, (b, e) in enumerate(self.bounds):
            s_run = self.s_runs[i & 1]
            with torch.cuda.stream(self.s_in):
                if trace is not None:
                    t_in0 = torch.cuda.Event(enable_timing=True); t_in0.record()
                self.d_start[b:e].copy_(h_start[b:e], non_blocking=True)
                ev_in = self.s_in.record_event(torch.cuda.Event(enable_timing=True) if trace is not None else None)
            with torch.cuda.stream(s_run):
                s_run.wait_event(ev_in)
                if trace is not None:
                    t_run0 = torch.cuda.Event(enable_timing=True); t_run0.record()
                f = self.d_out[0][b:e] if self.want_final else None
                L.check(lib.b2048_playout_fixed(ctx, L.ptr(self.d_start[b:e]), o, seed, gid_base + b, 0, L.ptr(f),
                                                L.ptr(self.d_out[1][b:e]), L.ptr(self.d_out[2][b:e]), e - b,
                                                C.c_void_p(s_run.cuda_stream)))
                ev_run = s_run.record_event(torch.cuda.Event(enable_timing=True) if trace is not None else None)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev_run)
                for h, d in zip(self.h_out, self.d_out):
                    if h is not None:
                        h[b:e].copy_(d[b:e], non_blocking=True)
                if trace is not None:
                    t_out1 = torch.cuda.Event(enable_timing=True); t_out1.record()
                    trace.append((t_in0, ev_in, t_run0, ev_run, t_out1))
        cur.wait_stream(self.s_out)
        self.s_out.synchronize()
        if trace is not None:
            self.last_events = trace
        return self.h_out


def set_playout_variant(variant, device=None):
    """K2 kernel variant: 2 (default) pair table (both slides of a row per lookup), 1 legality-first uniform slide,
    0 the L,U,D,R attempt cascade."""
    L.check(L.load_library().b2048_set_playout_variant(L.ctx(device), int(variant)))


def set_latency_mode(mode, device=None):
    """Small tensor-core launches (b2048_set_latency_mode): 0 = no latency variant, 1 = 16-warp one-set kernel,
    2 = + two-ply speculation for launches with few lines (<= 6 per SM), 3 (default) = + three-ply speculation for
    launches with at most 2 lines per SM.  Results are identical in every mode."""
    L.check(L.load_library().b2048_set_latency_mode(L.ctx(device), int(mode)))


def num_sms(device=None):
    """SM count of the context's device (persistent kernels launch one CTA per SM)."""
    return int(L.load_library().b2048_num_sms(L.ctx(device)))


def pipe_rate(kind, iters=20000, reps=3, device=None):
    """Measured SM integer pipe rate in thread-instructions per second (b2048_probe_pipe_rate; the denominators of the K2
    roofline): kind 0 = LOP3 only (ALU pipe), kind 1 = LOP3 + IMAD alternating (the issue limit).  Device timed, best of
    `reps` after one warm-up launch."""
    lib, ctx = L.load_library(), L.ctx(device)
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else torch.device(device).index or 0)
    sink = torch.zeros(1, dtype=torch.int32, device=dev)
    n = C.c_uint64(0)
    best = 0.0
    with torch.cuda.device(dev):
        for r in range(reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            L.check(lib.b2048_probe_pipe_rate(ctx, int(kind), int(iters), L.ptr(sink), C.byref(n), L.stream_ptr(dev)))
            e1.record()
            e1.synchronize()
            if r:
                best = max(best, n.value / (e0.elapsed_time(e1) * 1e-3))
    return best


def game_summary(finals, score, moves=None, acc=None, want_maxtile=False):
    """Accumulate the distribution summary of finished games (b2048_game_summary) into acc (int64[20], created
    zeroed when None).  -> (acc, maxtile u8[n] or None)"""
    d = _dev(finals)
    n = finals.numel()
    if acc is None:
        acc = torch.zeros(20, dtype=torch.int64, device=d)
    mt = torch.empty(n, dtype=torch.uint8, device=d) if want_maxtile else None
    L.check(L.load_library().b2048_game_summary(L.ctx(d), L.ptr(finals), L.ptr(score), L.ptr(moves), L.ptr(mt), L.ptr(acc), n,
                                                L.stream_ptr(d)))
    return acc, mt


def selfplay_fixed_games(G, number, seed, times=1, mode=2, starts=None, game_id_base=0, games_total=None, max_steps=4096,
                         record=True, device="cuda"):
    """Whole fixed-order root-search main games on the device (b2048_selfplay_fixed): one CTA per game.
    -> dict(boards i64[G,max_steps], results f32[G,max_steps,4] (record=True), final i64[G], score i64[G], steps i32[G],
            status i32[G] (1 = stopped at max_steps), total_moves i64[1])"""
    L.require_cuda()
    d = torch.device(device) if starts is None else _dev(starts)
    games_total = game_id_base + G if games_total is None else games_total
    r = dict(boards=torch.empty((G, max_steps), dtype=torch.int64, device=d) if record else None,
             results=torch.empty((G, max_steps, 4), dtype=torch.float32, device=d) if record else None,
             final=torch.empty(G, dtype=torch.int64, device=d), score=torch.empty(G, dtype=torch.int64, device=d),
             steps=torch.empty(G, dtype=torch.int32, device=d), status=torch.empty(G, dtype=torch.int32, device=d),
             total_moves=torch.zeros(1, dtype=torch.int64, device=d))
    L.check(L.load_library().b2048_selfplay_fixed(
        L.ctx(d), L.ptr(starts), G, game_id_base, games_total, number, times, mode, seed, max_steps, L.ptr(r["boards"]),
        L.ptr(r["results"]), L.ptr(r["final"]), L.ptr(r["score"]), L.ptr(r["steps"]), L.ptr(r["status"]),
        L.ptr(r["total_moves"]), L.stream_ptr(d)))
    return r


def selfplay_nn_games(blob, precision, G, number, seed, starts=None, game_id_base=0, games_total=None, max_steps=4096,
                      record=True, device="cuda"):
    """Whole NN-policy root-search main games on the device (b2048_selfplay_nn; selfplay.py:62-115 with mcts_nn as the
    search): one persistent launch, lines of all main games flow through a device-side queue, no per-step barrier.
    -> dict(boards i64[G,max_steps], results f32[G,max_steps,4] (record=True), final i64[G], score i64[G], steps i32[G],
            status i32[G] (1 = stopped at max_steps), total_moves i64[1])"""
    L.require_cuda()
    d = torch.device(device) if starts is None else _dev(starts)
    games_total = game_id_base + G if games_total is None else games_total
    lib = L.load_library()
    wbytes = lib.b2048_selfplay_nn_workspace_bytes(G, int(number))
    if wbytes < 0:
        raise ValueError("G * 4 * number is too large for one launch")
    work = torch.empty(wbytes, dtype=torch.uint8, device=d)  # torch's allocator returns 512-byte aligned blocks
    r = dict(boards=torch.empty((G, max_steps), dtype=torch.int64, device=d) if record else None,
             results=torch.empty((G, max_steps, 4), dtype=torch.float32, device=d) if record else None,
             final=torch.empty(G, dtype=torch.int64, device=d), score=torch.empty(G, dtype=torch.int64, device=d),
             steps=torch.empty(G, dtype=torch.int32, device=d), status=torch.empty(G, dtype=torch.int32, device=d),
             total_moves=torch.zeros(1, dtype=torch.int64, device=d), workspace=work)
    L.check(lib.b2048_selfplay_nn(
        L.ctx(d), PRECISIONS[precision], L.ptr(blob), L.ptr(starts), G, game_id_base, games_total, int(number), seed,
        int(max_steps), L.ptr(r["boards"]), L.ptr(r["results"]), L.ptr(r["final"]), L.ptr(r["score"]), L.ptr(r["steps"]),
        L.ptr(r["status"]), L.ptr(r["total_moves"]), L.ptr(work), L.stream_ptr(d)))
    return r
