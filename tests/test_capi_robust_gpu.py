"""GPU: re-entrancy of the C ABI -- concurrent launches from several host threads on several streams of one context
give the serial results (atomic work-cursor pool, no shared scratch), and every call runs on its context's device
without changing the caller's current device."""
import threading

import numpy as np
import pytest
import torch

from conftest import state_dict_from_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def net(pkg, ng):
    m = pkg.network.ConvNet(64, 3, precision="fp16")
    m.load_state_dict(state_dict_from_golden(ng, "random"))
    return m.cuda().eval()


def _jobs(pkg, net, ng):
    eng = pkg.engine
    blob = net.blob("fp16")
    origins = eng.init_boards(24, 5, 0)
    boards = torch.from_numpy(ng["lg_boards"][:3000].view(np.int64).copy()).cuda()

    def fixed(seed):
        f, s, c = eng.playout_fixed(30000, seed, 0)
        return torch.stack([f, s.to(torch.int64), c.to(torch.int64)])

    def root_fixed(seed):
        r = eng.mcts_fixed_lines(origins, 9, seed, 3, per_line=False, fused=True)
        return torch.cat([r["logsum"].flatten(), r["count"].flatten().to(torch.int64), r["minmov"].flatten().to(torch.int64)])

    def root_nn(seed):
        r = eng.mcts_nn_lines(blob, "fp16", origins, 7, seed, 11)
        return r["minmov"].flatten().to(torch.int64)

    def games_nn(seed):
        r = eng.playout_nn(blob, "fp16", 500, seed, 0)
        return torch.stack([r["final"], r["score"].to(torch.int64), r["moves"].to(torch.int64)])

    def forward(_seed):
        return net.forward_boards(boards, precision="fp16")

    def main_games(seed):
        r = eng.selfplay_nn_games(blob, "fp16", 3, 4, seed, max_steps=6)
        return torch.cat([r["final"], r["score"], r["steps"].to(torch.int64), r["boards"].flatten()])

    return [fixed, root_fixed, root_nn, games_nn, forward, main_games]


def test_concurrent_streams_and_threads_equal_serial(pkg, net, ng):
    jobs = _jobs(pkg, net, ng)
    plan = [(k % len(jobs), 100 + k) for k in range(36)]  # (job, seed)
    serial = []
    for j, seed in plan:
        serial.append(jobs[j](seed).clone())
    torch.cuda.synchronize()
    assert pkg._lib.take_error_flag() == 0
    results = [None] * len(plan)
    errors = []

    def worker(tid, nthreads):
        try:
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                for k in range(tid, len(plan), nthreads):
                    j, seed = plan[k]
                    results[k] = jobs[j](seed)
            st.synchronize()
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(t, 4)) for t in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    torch.cuda.synchronize()
    assert not errors, errors
    for k, (a, b) in enumerate(zip(serial, results)):
        assert torch.equal(a, b), ("launch %d (job %d) differs under concurrency" % (k, plan[k][0]))
    assert pkg._lib.take_error_flag() == 0


def test_launch_counter_counts_kernels(pkg):
    n0 = pkg._lib.launch_count()
    pkg.engine.playout_fixed(1000, 1, 0)
    assert pkg._lib.launch_count() - n0 == 1
    o = pkg.engine.init_boards(4, 1, 0)
    n1 = pkg._lib.launch_count()
    pkg.engine.mcts_fixed_lines(o, 5, 1, 0)
    assert pkg._lib.launch_count() - n1 == 2  # root moves + the playout kernel


def test_calls_keep_the_callers_current_device(pkg):
    import ctypes as C

    L = pkg._lib
    lib = L.load_library()
    cur = torch.cuda.current_device()
    h = C.c_void_p()
    L.check(lib.b2048_create(cur, C.byref(h)))
    assert torch.cuda.current_device() == cur
    L.check(lib.b2048_destroy(h))
    bad = C.c_void_p()
    assert lib.b2048_create(torch.cuda.device_count() + 3, C.byref(bad)) != 0
    assert b"no such CUDA device" in lib.b2048_last_error()
    assert torch.cuda.current_device() == cur


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_context_of_another_gpu_from_one_process(pkg):
    # a context of GPU 1 driven from a thread whose current device is GPU 0: same games, current device untouched
    assert torch.cuda.current_device() == 0
    a = pkg.engine.playout_fixed(5000, 77, 0, device="cuda:0")
    b = pkg.engine.playout_fixed(5000, 77, 0, device="cuda:1")
    assert torch.cuda.current_device() == 0
    torch.cuda.synchronize(0)
    torch.cuda.synchronize(1)
    for x, y in zip(a, b):
        assert x.device.index == 0 and y.device.index == 1 and torch.equal(x.cpu(), y.cpu())
    o1 = pkg.engine.init_boards(6, 3, 0, device="cuda:1")
    r1 = pkg.engine.mcts_fixed_lines(o1, 5, 3, 0)
    r0 = pkg.engine.mcts_fixed_lines(o1.to("cuda:0"), 5, 3, 0)
    assert torch.equal(r1["lmoves"].cpu(), r0["lmoves"].cpu()) and torch.cuda.current_device() == 0


def test_latency_modes_give_identical_results(pkg, net, ng):
    # mode 0: two tile-sets / 8-warp groups only; 1: 16-warp one-set kernel for small launches; 2: + two-ply speculation
    # (a line's board and its four successors in one forward); 3: + three-ply speculation (up to 11 second successors as
    # well).  Same boards evaluated, same logits, same moves: every result must be bit-identical.  Sizes chosen so that
    # modes 2 and 3 really speculate (<= 6 x #SMs resp. <= 2 x #SMs lines in flight).
    eng = pkg.engine
    blob = net.blob("fp16")
    origins = eng.init_boards(3, 9, 0)
    boards = torch.from_numpy(ng["lg_boards"][:1000].view(np.int64).copy()).cuda()
    ref = None
    try:
        for mode in (0, 1, 2, 3):
            eng.set_latency_mode(mode)
            r = eng.mcts_nn_lines(blob, "fp16", origins, 20, 77, 5, per_line=False)
            g = eng.playout_nn(blob, "fp16", 100, 78, 0)
            grp = eng.playout_nn(blob, "fp16", 120, 79, 0, group_size=30)
            sp = eng.selfplay_nn_games(blob, "fp16", 2, 10, 80, max_steps=2048)
            # whole-game launches too big to speculate from the start: modes >= 2 regroup the survivors of every tile in the
            # tail of the launch (one-set kernel at 3,000 games, two tile-sets at 12,000)
            # the precise tensor-core mode goes through the same latency regimes
            blob2 = net.blob("fp16x2")
            r2 = eng.mcts_nn_lines(blob2, "fp16x2", origins, 20, 77, 5, per_line=False)
            g2 = eng.playout_nn(blob2, "fp16x2", 100, 78, 0)
            sp2 = eng.selfplay_nn_games(blob2, "fp16x2", 2, 10, 80, max_steps=40)
            big3 = eng.playout_nn(blob2, "fp16x2", 1500, 83, 0)
            big1 = eng.playout_nn(blob, "fp16", 3000, 81, 0)
            big2 = eng.playout_nn(blob, "fp16", 12000, 82, 0, group_size=40)
            fw = net.forward_boards(boards, precision="fp16")
            torch.cuda.synchronize()
            n0, n1 = int(sp["steps"][0]), int(sp["steps"][1])
            got = [r["minmov"], g["final"], g["score"], g["moves"], grp["group_min"], sp["final"], sp["score"], sp["steps"],
                   sp["boards"][0, :n0], sp["boards"][1, :n1], sp["results"][0, :n0], fw,
                   big1["final"], big1["score"], big1["moves"], big2["group_min"],
                   r2["minmov"], g2["final"], g2["moves"], sp2["boards"], sp2["results"], sp2["steps"], big3["final"], big3["moves"]]
            assert int(g["total_moves"].item()) == int(g["moves"].sum().item())
            assert n0 > 20 and n1 > 20 and int(sp["status"].sum()) == 0
            if ref is None:
                ref = [t.clone() for t in got]
            else:
                for k, (a, b) in enumerate(zip(ref, got)):
                    assert torch.equal(a, b), (mode, k)
        assert pkg._lib.take_error_flag() == 0
    finally:
        eng.set_latency_mode(3)
