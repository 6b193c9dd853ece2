"""CPU (gloo, world_size 2): the multi-GPU host logic -- line sharding and the root-statistics
allreduce -- reproduces the unsharded result bit-for-bit.  Per-rank partials come from the oracle
(the CUDA kernels need a GPU; their sharding invariance is checked in test_engine_gpu.py)."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _partial(oracle, origins, number, seed, begin, end, eval0=0):
    G = len(origins)
    legal = np.zeros((G, 4), np.uint8)
    logsum = np.zeros((G, 4), np.int64)
    count = np.zeros((G, 4), np.int32)
    minmov = np.full((G, 4), 2**31 - 1, np.int32)
    for g, o in enumerate(origins):
        lg, rs, sc, mv, _ = oracle.root_lines(int(o), number, seed, eval0 + g)
        legal[g] = lg
        for i in range(4):
            if lg[i] and end > begin:
                t = np.log10(sc[i, begin:end].astype(np.float64) + float(rs[i]) + 1.0)
                logsum[g, i] = np.rint(np.ldexp(t, 40)).astype(np.int64).sum()
                count[g, i] = end - begin
                minmov[g, i] = mv[i, begin:end].min()
    return legal, logsum, count, minmov


def _worker(rank, world_size, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    from oracle import oracle

    d = importlib.import_module("2048nn_b200.dist")
    origins = [0x2141348, 0x0000000100000001, 0x1111222233334444, 0x21]
    number, seed = 7, 4242
    b, e = d.shard_lines(number, rank, world_size)
    legal, logsum, count, minmov = _partial(oracle, origins, number, seed, b, e)
    stats = dict(logsum=torch.from_numpy(logsum), count=torch.from_numpy(count), minmov=torch.from_numpy(minmov))
    d.allreduce_root_stats(stats)
    mean = d.mean_log_from_stats(torch.from_numpy(legal), stats["logsum"], stats["count"])
    mn = d.min_from_stats(torch.from_numpy(legal), stats["minmov"])
    q.put((rank, (b, e), stats["logsum"].numpy().copy(), stats["count"].numpy().copy(), mean.numpy().copy(), mn.numpy().copy()))
    # origin sharding (G >= world size): the library's own driver with the oracle standing in for the kernels
    def run(o, eval0, lb, le):
        lg, ls, ct, mm = _partial(oracle, [int(x) for x in o.numpy().view(np.uint64)], number, seed, lb, le, eval0)
        return dict(legal=torch.from_numpy(lg), logsum=torch.from_numpy(ls), count=torch.from_numpy(ct), minmov=torch.from_numpy(mm),
                    total_moves=torch.tensor([int(le - lb) * len(o)]))
    t_orig = torch.from_numpy(np.array(origins, np.uint64).view(np.int64))
    plan = d.shard_plan(len(origins), number, rank, world_size)
    r = d._sharded(run, t_orig, number, 0, None)
    q.put((10 + rank, plan, r["logsum"].numpy().copy(), r["count"].numpy().copy(), r["minmov"].numpy().copy(), int(r["total_moves"])))
    # pipelined evaluations: three exchanges enqueued (async_op) before the first is waited for -- same statistics
    pend = [d._sharded(run, t_orig, number, 0, None, async_reduce=True) for _ in range(3)]
    same = True
    for pr in pend:
        assert "_pending" in pr
        d.wait_root_stats(pr)
        same &= bool(torch.equal(pr["logsum"], r["logsum"]) and torch.equal(pr["count"], r["count"]) and torch.equal(pr["minmov"], r["minmov"]))
    q.put((20 + rank, same))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_lines_partition():
    d = importlib.import_module("2048nn_b200.dist")
    for number in (1, 7, 10, 50):
        for ws in (1, 2, 3, 8):
            parts = [d.shard_lines(number, r, ws) for r in range(ws)]
            assert parts[0][0] == 0 and parts[-1][1] == number
            assert all(parts[i][1] == parts[i + 1][0] for i in range(ws - 1))
            sizes = [e - b for b, e in parts]
            assert max(sizes) - min(sizes) <= 1


def test_allreduce_root_stats_world2(orc):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(6)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got.sort(key=lambda t: t[0])
    out, by_origin, piped = got[:2], got[2:4], got[4:]
    assert all(p[1] for p in piped)  # async (pipelined) exchanges give the statistics of the blocking ones
    assert out[0][1] == (0, 4) and out[1][1] == (4, 7)
    # every rank holds the same reduced statistics ...
    for k in range(2, 6):
        assert np.array_equal(out[0][k], out[1][k])
    # ... equal to the unsharded evaluation
    origins = [0x2141348, 0x0000000100000001, 0x1111222233334444, 0x21]
    legal, logsum, count, minmov = _partial(orc, origins, 7, 4242, 0, 7)
    assert np.array_equal(out[0][2], logsum) and np.array_equal(out[0][3], count)
    for g, o in enumerate(origins):
        want_min = orc.mcts_fixed_min(o, 7, 4242, g)
        assert out[0][5][g].tolist() == want_min
        want_mean = orc.mcts_fixed(o, 7, 4242, g)
        assert np.allclose(out[0][4][g], [float(x) for x in want_mean], atol=1e-10, rtol=0)
    # origin-sharded evaluation: ranks own origins [0,2) and [2,4), all 7 lines, same reduced statistics
    assert by_origin[0][1] == (0, 2, 0, 7) and by_origin[1][1] == (2, 4, 0, 7)
    for o in by_origin:
        assert np.array_equal(o[2], logsum) and np.array_equal(o[3], count) and np.array_equal(o[4], minmov)
        assert o[5] == 7 * 2  # this rank's own lines: accounting is summed over ranks once per run, not per evaluation


def test_shard_plan():
    d = importlib.import_module("2048nn_b200.dist")
    assert d.shard_plan(1, 50, 3, 8) == (0, 1, 0, 50)                         # config 3: 200 lines -> replicated, no exchange
    assert d.shard_plan(2, 2000, 3, 8) == (0, 2) + d.shard_lines(2000, 3, 8)  # few origins, many lines -> lines
    assert d.shard_plan(256, 50, 3, 8) == (96, 128, 0, 50)                    # config 4: whole groups per rank
    for G, ws in ((8, 8), (9, 4), (2048, 8)):
        parts = [d.shard_plan(G, 10, r, ws) for r in range(ws)]
        assert parts[0][0] == 0 and parts[-1][1] == G and all(parts[i][1] == parts[i + 1][0] for i in range(ws - 1))


def test_shard_games_partition():
    d = importlib.import_module("2048nn_b200.dist")
    for G, ws in ((256, 8), (256, 1), (5, 2), (3, 8), (1, 4), (17, 4)):
        parts = [d.shard_games(G, r, ws) for r in range(ws)]
        covered = []
        for g0, n, per in parts:
            assert per == parts[0][2] and 0 <= n <= per
            covered += list(range(g0, g0 + n))
        assert covered == list(range(G))            # disjoint, ordered, complete
        assert ws * parts[0][2] >= G                # the gather buffer of `per` games per rank holds every record
    assert d.shard_games(256, 3, 8) == (96, 32, 32)  # BASELINE config 4 on 8 GPUs: 32 main games = 6,400 lines per GPU


def _epoch_worker(rank, world_size, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    tr = importlib.import_module("2048nn_b200.trainer")

    class Recorder:  # stands in for TrainStep (the CUDA step needs a GPU): remembers which records it was given
        def __init__(self):
            self.seen = []

        def step(self, boards, targets, group=None):
            self.seen.append(boards.clone())
            return torch.tensor(float(boards.numel()))

    torch.manual_seed(100 + rank)  # ranks with DIFFERENT generator states must still use one permutation
    n, bs = 53, 8
    boards = torch.arange(n, dtype=torch.int64)
    rec = Recorder()
    mean_loss, batches = tr.train_epoch(rec, boards, torch.zeros(n, 4), bs)
    q.put((rank, batches, mean_loss, torch.cat(rec.seen).tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_epoch_world2():
    # trainer.train_epoch under a process group: one permutation for all ranks, disjoint slices of every global batch,
    # a ragged last global batch that leaves a rank short is dropped on every rank (same number of steps everywhere)
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_epoch_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, b0, l0, s0), (_, b1, l1, s1) = got
    assert b0 == b1 == 3                      # 53 records, global batch 16: 3 full global batches, the ragged rest dropped
    assert len(s0) == len(s1) == 24 and not set(s0) & set(s1) and len(set(s0) | set(s1)) == 48
    assert l0 == l1 == 8.0
