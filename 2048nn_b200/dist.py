"""Multi-GPU sharding of root evaluations: one process per GPU, torch.distributed over NCCL.

Playouts are independent, so a root evaluation shards over the ranks with the same (seed, eval_id) keyed
streams -- by ORIGINS when there are at least as many origins as ranks (config 4: every rank evaluates whole
(origin, root move) groups, so the early stop of the min-move search prunes exactly as on one GPU), else by LINES
(rank r plays lines [begin_r, end_r) of every group; a config-3-sized evaluation that fits one GPU's game slots is
replicated instead, see shard_plan) -- and the ranks combine their partial statistics with ONE small allreduce
(SURVEY.md 8e):

    mean-log (mcts.py:4-34)      logsum i64[G,4] fixed-point + count i32[G,4]   SUM
    min-move (mcts.py:73-103,
              mcts_nn.py:4-43)   minmov i32[G,4]                                MIN

Because the streams are keyed by (seed, game id) and the fused sums are integers, the reduced
statistics are bit-identical for every world size (tests/test_dist_cpu.py, test_engine_gpu.py).
The reduction helpers take plain tensors, so the host logic is testable with gloo on CPU.
"""
import numpy as np
import torch
import torch.distributed as dist

LOGSUM_SCALE = float(2 ** 40)  # B2048_LOGSUM_FRAC_BITS
INT32_MAX = 2 ** 31 - 1


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_lines(number, rank, world_size):
    """Contiguous split of `number` lines over ranks: sizes differ by at most one."""
    base, rem = divmod(int(number), int(world_size))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_games(G, rank, world_size):
    """Device-resident main games (selfplay_batch) shard by game id: rank r plays games [g0, g0 + n) of the session;
    every rank's share is padded to `per` games when the records are gathered.  -> (g0, n, per)"""
    per = (int(G) + int(world_size) - 1) // int(world_size)
    g0 = min(int(G), int(rank) * per)
    return g0, max(0, min(int(G), g0 + per) - g0), per


REPLICATE_BELOW_LINES = 148 * 32  # one 32-board tile-set per SM: a root evaluation this small is latency-bound


def shard_plan(G, number, rank, world_size):
    """(origin_begin, origin_end, line_begin, line_end) of this rank's share of a G x 4 x number root evaluation.
    The whole range on every rank means "replicated": fewer origins than ranks and so few lines that one GPU runs
    them all concurrently -- the call is bound by the longest line, splitting it only weakens the early stop and
    adds an exchange, so every rank computes the (identical, deterministic) result itself and nothing is reduced."""
    if G >= world_size:
        o0, o1 = shard_lines(G, rank, world_size)
        return o0, o1, 0, int(number)
    if world_size > 1 and 4 * int(G) * int(number) <= REPLICATE_BELOW_LINES:
        return 0, int(G), 0, int(number)
    l0, l1 = shard_lines(number, rank, world_size)
    return 0, int(G), l0, l1


def embed_partial(part, G, o0, o1):
    """Partial statistics of origins [o0, o1) -> full [G,4] tensors that are neutral elsewhere (0 for the SUM
    statistics, INT32_MAX for minmov), ready for allreduce_root_stats."""
    out = {}
    for k, neutral in (("logsum", 0), ("count", 0), ("minmov", INT32_MAX)):
        if part.get(k) is not None:
            full = torch.full((G, 4), neutral, dtype=part[k].dtype, device=part[k].device)
            if o1 > o0:
                full[o0:o1] = part[k]
            out[k] = full
    return out


def allreduce_root_stats(stats, group=None, async_op=False):
    """In-place allreduce of a partial-statistics dict (keys present among logsum/count: SUM;
    minmov: MIN).  No-op without an initialised process group.

    async_op=True only ENQUEUES the collectives (they run on the communication stream once this rank's kernel is
    done) and returns at once, so the caller's stream can start the next root evaluation without waiting for the
    slowest rank: the statistics are valid after wait_root_stats(stats).  Independent evaluations pipelined this way
    pay the load imbalance between ranks once per pipeline instead of once per evaluation."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return stats
    pending = []
    if "logsum" in stats and stats["logsum"] is not None:
        packed = torch.stack([stats["logsum"], stats["count"].to(torch.int64)])  # one collective for both
        work = dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        pending.append((work, packed))
    if "minmov" in stats and stats["minmov"] is not None:
        pending.append((dist.all_reduce(stats["minmov"], op=dist.ReduceOp.MIN, group=group, async_op=async_op), None))
    stats["_pending"] = pending
    return stats if async_op else wait_root_stats(stats)


def wait_root_stats(stats):
    """Complete the collectives enqueued by allreduce_root_stats (the current stream waits for them) and unpack."""
    for work, packed in stats.pop("_pending", []):
        if work is not None:
            work.wait()
        if packed is not None:
            stats["logsum"].copy_(packed[0])
            stats["count"].copy_(packed[1].to(stats["count"].dtype))
    return stats


def mean_log_from_stats(legal, logsum, count):
    """mcts_fixed values from reduced statistics: mean log10 per root move, -1 where illegal."""
    legal = legal.to(torch.bool)
    mean = logsum.to(torch.float64) / LOGSUM_SCALE / count.clamp(min=1).to(torch.float64)
    return torch.where(legal, mean, torch.full_like(mean, -1.0))


def min_from_stats(legal, minmov, illegal_value=0):
    """mcts_fixed_min / mcts_nn values: 1 + min_j L_j, `illegal_value` where the move is illegal."""
    legal = legal.to(torch.bool)
    return torch.where(legal, minmov.to(torch.int64) + 1, torch.full_like(minmov, illegal_value, dtype=torch.int64))


# ---- sharded root evaluations on the GPU --------------------------------------------------------
def _sharded(run, origins, number, eval_id_base, group, async_reduce=False):
    """run(origins, eval_id_base, line_begin, line_end) -> statistics dict of those origins / lines."""
    rank, ws = world()
    G = origins.numel()
    o0, o1, l0, l1 = shard_plan(G, number, rank, ws)
    if ws > 1 and (o0, o1, l0, l1) == (0, G, 0, int(number)):
        r = run(origins, eval_id_base, 0, int(number))  # replicated: identical on every rank, no exchange
        if rank and r.get("total_moves") is not None:
            r["total_moves"].zero_()  # count the replicated work once
        return r
    if (o0, o1) == (0, G):
        r = run(origins, eval_id_base, l0, l1)
    else:
        r = run(origins, eval_id_base, 0, 0)  # root moves of every origin (legal, rscore); no lines
        part = run(origins[o0:o1].contiguous(), eval_id_base + o0, l0, l1)  # G >= world size: o1 > o0 on every rank
        r.update(embed_partial(part, G, o0, o1))
        if part.get("total_moves") is not None:
            r["total_moves"] = part["total_moves"]
    allreduce_root_stats(r, group, async_op=async_reduce)
    # r["total_moves"] (throughput accounting) stays THIS RANK's count: callers sum it over ranks once, at the end of
    # a run, instead of paying a second collective per root evaluation (sum_over_ranks below)
    return r


def sum_over_ranks(t, group=None):
    """In-place SUM allreduce of an accounting tensor (no-op without a process group)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def mcts_fixed_sharded(origins, number, seed, eval_id_base=0, group=None):
    """Fixed-order root search for G origins sharded over the process group (shard_plan).
    Returns dict(legal, rscore, logsum, count, minmov, mean, min) -- identical on every rank."""
    from . import engine

    r = _sharded(lambda o, e, b, l: engine.mcts_fixed_lines(o, number, seed, e, b, l, per_line=False, fused=True),
                 origins, number, eval_id_base, group)
    r["mean"] = mean_log_from_stats(r["legal"], r["logsum"], r["count"])
    r["min"] = min_from_stats(r["legal"], r["minmov"])
    return r


def mcts_nn_sharded(blob, precision, origins, number, seed, eval_id_base=0, group=None, async_reduce=False):
    """NN-policy root search (mcts_nn.py:4-43) sharded over the process group (shard_plan).  With line sharding the
    early stop uses each rank's local minimum; the allreduce(MIN) restores the global one.
    async_reduce=True returns as soon as this rank's launch and the collective are enqueued; call
    mcts_nn_finish(r) before reading r["minmov"] / r["min"] (pipelining of independent evaluations)."""
    from . import engine

    r = _sharded(lambda o, e, b, l: engine.mcts_nn_lines(blob, precision, o, number, seed, e, b, l),
                 origins, number, eval_id_base, group, async_reduce)
    return r if async_reduce else mcts_nn_finish(r)


def mcts_nn_finish(r):
    """Second half of mcts_nn_sharded(async_reduce=True): wait for the exchange, derive the values."""
    wait_root_stats(r)
    r["min"] = min_from_stats(r["legal"], r["minmov"])
    return r
