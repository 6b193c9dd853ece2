#!/usr/bin/env python
"""bench.py -- playout moves/sec of the B200 engine (BASELINE.json metric), one JSON line.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--games G] [--impl ours|reference]

N > 1 is launched by the driver through torch.distributed.run (one rank per GPU, NCCL).
A "step" is one pass of the hot path (play_fixed_batch: every game played to termination
inside one persistent kernel) over one batch of `--games` synthetic fresh start boards per GPU.
Per-GPU work is fixed as N grows ("weak"); the path shards by game with no data-path
collective -- ranks only exchange their move counts and times for the report.

Blocks of the line besides the contract's keys: `roofline.executed` (thread-instructions per move and pipe utilisation from the
committed ncu capture, and `measured_pipe_rates`: this GPU's LOP3 / LOP3+IMAD rates measured in the run); `nn_policy` = the second
half of the metric (K pipelined mcts_nn root evaluations, 256 origins per GPU, with `e2e`, `roofline`, `cpu_baseline`,
`forward_only` + `torch_baseline` (config 5 point and the same-GPU PyTorch arms) and `precise_mode`: the tensor-core mode that
keeps every board within 1e-2 of fp32 -- forward, root evaluation, the same K-evaluation loop over all ranks, config 3);
`mcts_nn_game` = config 3 (one main game, ms per mcts_nn(number=50) main move); `selfplay256` = config 4 (256 main games in
total, strong scaling); `train_step`; `mcts_fixed_game` = config 1.

`--impl reference` times the CPU implementation of the same path on the box's host cores.  The
reference is pure Python (no compiled part), so the arm runs the C restatement in oracle/
(kind "port") on all host threads -- a faster baseline than the Python original.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "playout moves/sec (fixed-order, play_fixed_batch to termination)"
NN_METRIC = "playout moves/sec (c64b3 NN-policy, mcts_nn root evaluations)"
UNIT = "moves/s"
INSTR_PER_MOVE = 250.0  # algorithmic thread-instructions per accepted move


_T0 = time.perf_counter()


def log(msg):
    """progress on stderr (stdout carries exactly the one JSON line)"""
    if os.environ.get("RANK", "0") == "0":
        sys.stderr.write("[bench %7.2fs] %s\n" % (time.perf_counter() - _T0, msg))
        sys.stderr.flush()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        hi = [x for x in sm if mx and x > 0.5 * mx] or sm  # samples under load
        med = hi[len(hi) // 2] if hi else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_rate(target_seconds, nthreads):
    """Oracle port (C) on the host: play_fixed_batch on a bounded sample -> (moves/s, sample str)."""
    from oracle import oracle

    oracle.lib()
    t0 = time.perf_counter()
    _, _, c = oracle.play_fixed_batch(20000, 99, 0, nthreads=nthreads)
    dt = time.perf_counter() - t0
    rate = float(c.sum()) / dt
    n = int(max(20000, min(20_000_000, rate * target_seconds / 225.0)))
    t0 = time.perf_counter()
    _, _, c = oracle.play_fixed_batch(n, 100, 0, nthreads=nthreads)
    dt = time.perf_counter() - t0
    return float(c.sum()) / dt, f"{n} fresh fixed-order games ({int(c.sum())} moves) in {dt:.2f}s wall"


def shipped_state_dict():
    """The c64b3 checkpoint of the reference (models/20200213, fixture copy) as a state dict, or None."""
    import numpy as np
    import torch

    gpath = os.path.join(ROOT, "tests", "golden", "c64b3_golden.npz")
    if not os.path.exists(gpath):
        return None
    ng = np.load(gpath)
    return {k[len("sd_shipped/"):]: torch.from_numpy(np.array(ng[k])) for k in ng.files if k.startswith("sd_shipped/")}


def cpu_nn_rate(target_seconds, nthreads):
    """Oracle port of mcts_nn.mcts_nn(model, b, 50, 'cpu') (mcts_nn.py:4-43: torch fp32 ConvNet on `nthreads` intra-op
    threads + the C board engine) on fresh boards -> (NN-policy moves/s, ms per call, sample str)."""
    import torch
    from oracle import oracle

    oracle.lib()
    torch.set_num_threads(nthreads)
    sd = shipped_state_dict()
    calls, moves, t0 = 0, 0, time.perf_counter()
    while True:
        b = oracle.init_board(oracle.rng_key(4242, calls))
        _, mv = oracle.mcts_nn(sd, b, 50, 100, calls, want_moves=True)
        calls += 1
        moves += mv
        dt = time.perf_counter() - t0
        if dt >= target_seconds or calls >= 8:
            break
    return moves / dt, dt / calls * 1e3, f"{calls} mcts_nn(number=50) call(s) from fresh boards, shipped c64b3, {moves} policy moves in {dt:.2f}s wall"


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import oracle

    cores = oracle.lib().orc_num_threads()
    workload = workload_name(args.games)
    for _ in range(args.warmup):
        cpu_port_rate(0.3, cores)
    rates, sample = [], ""
    t_all = time.perf_counter()
    for _ in range(args.steps):
        r, sample = cpu_port_rate(max(1.0, min(8.0, 60.0 / max(1, args.steps))), cores)
        rates.append(r)
    wall = time.perf_counter() - t_all
    v = sum(rates) / len(rates)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload, "note": "reference is pure Python; timed here as its C restatement "
                   "(oracle/b2048_oracle.c) on all host threads, each step a bounded sample of the workload"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if not args.no_nn:
        nv, ms_call, nsample = cpu_nn_rate(args.cpu_seconds, cores)
        line["nn_policy"] = {"metric": NN_METRIC, "value": nv, "unit": UNIT, "ms_per_mcts_nn_call": ms_call,
                             "cpu_baseline": {"value": nv, "unit": UNIT, "cores": cores, "kind": "port", "sample": nsample}}
    print(json.dumps(line), flush=True)


def torch_forward(m, x):
    """The reference's ConvNet.forward (network.py:78-87) in plain PyTorch on the module tree of the drop-in
    (cuDNN / cuBLAS kernels): the existing Blackwell implementation the hand-written forward has to beat."""
    import torch

    x = m.in_block(x)
    for blk in m.blocks:
        x = torch.relu(x + blk(x))
    x = m.out_block(x)
    return torch.log_softmax(m.policy(x.reshape(-1, m.out_size)), dim=1)


def torch_forward_baselines(pkg, m, boards):
    """Eager fp32 and channels_last + bf16-autocast forward of the same network on the same GPU, pre-encoded one-hot
    input already resident (the encoder is not timed: favourable to the baseline)."""
    import torch

    n, chunk = boards.numel(), 1 << 17
    out = {}
    with torch.no_grad():
        x = pkg.engine.encode_boards(boards[:chunk].contiguous())  # (chunk,16,4,4) fp32
        for name in ("eager_fp32", "channels_last_bf16_autocast"):
            if name == "eager_fp32":
                xin, mod = x, m

                def run():
                    return torch_forward(mod, xin)
            else:
                xin, mod = x.contiguous(memory_format=torch.channels_last), m.to(memory_format=torch.channels_last)

                def run():
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        return torch_forward(mod, xin)
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = max(1, n // chunk)
            e0.record()
            for _ in range(reps):
                run()
            e1.record()
            torch.cuda.synchronize()
            out[name] = reps * chunk / (e0.elapsed_time(e1) * 1e-3)
    out["note"] = ("same module tree through cuDNN/cuBLAS, %d boards as %d launches-sets of %d pre-encoded one-hot boards "
                   "(1,024 B/board resident in HBM; encoder not timed)" % (n, max(1, n // chunk), chunk))
    return out


def main_games_block(pkg, blob, G_total, number, main_steps, steps, warmup, seed0, rank, world, barrier):
    """Device-resident NN main games (b2048_selfplay_nn): G_total main games of a session sharded over the ranks by
    game id (strong scaling), each played for `main_steps` main moves per bench step.  Timed twice: device-resident
    (start boards on the device, records left on the device) and end to end (pinned host start boards in, records
    and results copied back to pinned host memory inside the timed region)."""
    import torch

    eng = pkg.engine
    g0, n_local, _ = importlib.import_module("2048nn_b200.dist").shard_games(G_total, rank, world)
    lc = pkg._lib.launch_count

    def play(seed, starts=None):
        if not n_local:
            return None
        return eng.selfplay_nn_games(blob, "fp16", n_local, number, seed, starts=starts, game_id_base=g0, games_total=G_total,
                                     max_steps=main_steps)

    log("main games: %d total, %d here, %d main moves per step: warm-up" % (G_total, n_local, main_steps))
    for w in range(max(3, warmup)):
        play(seed0 + w)
    barrier()
    log("main games: timed region")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    moves = torch.zeros(1, dtype=torch.int64, device="cuda")
    mains = torch.zeros(1, dtype=torch.int64, device="cuda")
    l0 = lc()
    e0.record()
    for k in range(steps):
        r = play(seed0 + 100 + k)
        if r is not None:
            moves += r["total_moves"]
            mains += r["steps"].sum()
    e1.record()
    barrier()
    launches = lc() - l0
    # end to end: host start boards -> device, main games, records -> host
    h2d = d2h = 0
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_moves = 0
    if n_local:
        h_start = eng.init_boards(n_local, seed0 + 999, g0).cpu().pin_memory()
        d_start = torch.empty(n_local, dtype=torch.int64, device="cuda")
        hb = torch.empty((n_local, main_steps), dtype=torch.int64).pin_memory()
        hr = torch.empty((n_local, main_steps, 4), dtype=torch.float32).pin_memory()
        hs = torch.empty((3, n_local), dtype=torch.int64).pin_memory()
        hm = torch.empty(1, dtype=torch.int64).pin_memory()
        h2d = n_local * 8
        d2h = hb.numel() * 8 + hr.numel() * 4 + hs.numel() * 8 + 8
    def e2e_step(seed):
        d_start.copy_(h_start, non_blocking=True)
        r = play(seed, starts=d_start)
        hb.copy_(r["boards"], non_blocking=True)
        hr.copy_(r["results"], non_blocking=True)
        hs.copy_(torch.stack([r["final"], r["score"], r["steps"].to(torch.int64)]), non_blocking=True)
        hm.copy_(r["total_moves"], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return int(hm[0]) + 0 * int(hs[1, 0])  # the step's results, read on the host

    if n_local:
        e2e_step(seed0 + 199)  # untimed: the first use of the copy / stack / cast kernels loads them
    barrier()
    log("main games: end-to-end region")
    t0.record()
    for k in range(steps):
        if n_local:
            e2e_moves += e2e_step(seed0 + 200 + k)
    t1.record()
    barrier()
    return {"ms": e0.elapsed_time(e1), "e2e_ms": t0.elapsed_time(t1), "moves": float(moves.item()), "mains": float(mains.item()),
            "e2e_moves": float(e2e_moves), "launches": launches, "h2d": h2d, "d2h": d2h, "games_local": n_local}


def nn_policy_block(pkg, args, rank, world, barrier):
    """K steps of one sharded mcts_nn root evaluation: G origins x 4 root moves x `number` lines, sharded over the
    ranks by dist.shard_plan (whole origins per rank when G >= ranks), [G,4] MIN allreduce (NCCL) per step."""
    import numpy as np
    import torch

    d = importlib.import_module("2048nn_b200.dist")
    eng = pkg.engine
    m = pkg.network.ConvNet(channels=64, blocks=3, precision="fp16")
    weights = "random-init (torch.manual_seed(0))"
    sd = shipped_state_dict()
    if sd is not None:
        m.load_state_dict(sd)
        weights = "models/20200213 c64b3 checkpoint (fixture copy)"
    m.cuda().eval()
    blob = m.blob("fp16")
    lc = pkg._lib.launch_count
    G, number = args.nn_origins * world, args.nn_lines
    origins = eng.init_boards(G, 4242, 0)  # synthetic fresh start boards, same on every rank
    log("nn_policy: root evaluations")

    def timed_evaluations(blob, prec, steps, warm, seed0):
        """`steps` pipelined root evaluations of all G origins (sharded by dist.shard_plan), device timed; -> (ms, moves of this
        rank as a device tensor, launches)"""
        cur = torch.cuda.current_stream()
        lanes = [torch.cuda.Stream() for _ in range(max(2, args.nn_lanes))] if not args.nn_one_stream else [cur]
        for w in range(warm):  # warm-up on the lanes themselves (the first launch on a fresh stream was seen to cost ~0.1 s once)
            lanes[w % len(lanes)].wait_stream(cur)
            with torch.cuda.stream(lanes[w % len(lanes)]):
                d.mcts_nn_sharded(blob, prec, origins, number, seed0 + w, 0)
        for st in lanes:
            cur.wait_stream(st)
        torch.cuda.synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        moves = torch.zeros(1, dtype=torch.int64, device="cuda")
        l0 = lc()
        e0.record()
        for st in lanes:
            st.wait_stream(cur)
        part = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in lanes]
        inflight = []
        for k in range(steps):
            i = k % len(lanes)
            with torch.cuda.stream(lanes[i]):
                r = d.mcts_nn_sharded(blob, prec, origins, number, seed0 + 100 + k, k * G, async_reduce=True)
                part[i] += r["total_moves"]  # this rank's lines; summed over ranks once, after the timed region
                inflight.append((i, r))
        for i, r in inflight:
            with torch.cuda.stream(lanes[i]):
                d.mcts_nn_finish(r)
        for st in lanes:
            cur.wait_stream(st)
        for p_ in part:
            moves += p_
        e1.record()
        barrier()
        return e0.elapsed_time(e1), moves, lc() - l0

    # Independent evaluations are pipelined twice over: (1) launches alternate between two streams, so the persistent CTAs of
    # evaluation k+1 take over the SMs that evaluation k's CTAs leave while its last lines are still running (the library is
    # re-entrant per stream); (2) the [G,4] MIN exchange of evaluation k is only enqueued (dist.allreduce_root_stats async_op),
    # so a rank does not wait for the slowest rank at every step.  Every launch and exchange completes inside the timed region.
    ms, moves, launches = timed_evaluations(blob, "fp16", args.steps, max(3, args.warmup), 100)
    # end to end through the reference-facing call shape: host origins in, host [G,4] results out
    h_orig = origins.cpu().pin_memory()
    h_res = torch.empty((G, 4), dtype=torch.int64).pin_memory()
    d_orig = torch.empty_like(origins)
    e2e_moves = torch.zeros(1, dtype=torch.int64, device="cuda")
    def e2e_step(k):
        d_orig.copy_(h_orig, non_blocking=True)
        r = d.mcts_nn_sharded(blob, "fp16", d_orig, number, 300 + k, k * G)
        h_res.copy_(r["min"], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        _ = int(h_res[0, 0])  # host read of the step's result
        return r["total_moves"]

    e2e_step(args.steps)  # untimed: first use of the copies
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for k in range(args.steps):
        e2e_moves += e2e_step(k)
    t1.record()
    barrier()
    d.sum_over_ranks(moves)
    d.sum_over_ranks(e2e_moves)
    plan = d.shard_plan(G, number, rank, world)
    how = ("replicated on every rank" if world > 1 and plan == (0, G, 0, number) else
           "whole origins per rank (%d each)" % (plan[1] - plan[0]) if plan[2:] == (0, number) and world > 1 else
           "lines of every root move split over the ranks" if world > 1 else "one GPU")
    # config 5: forward-only sweep point (1M boards per GPU from policy-game positions), device timed
    log("nn_policy: forward only + torch baselines")
    nb = 1 << 20
    fb = eng.init_boards(nb, 99, rank * nb)
    for _ in range(3):
        m.forward_boards(fb, precision="fp16")
    torch.cuda.synchronize()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(3):
        m.forward_boards(fb, precision="fp16")
    f1.record()
    torch.cuda.synchronize()
    fwd_boards_per_s = 3 * nb / (f0.elapsed_time(f1) * 1e-3)
    # the precise tensor-core mode (fp16 hi+lo operands on L1..L3: every board within 1e-2): its cost, same workloads
    for _ in range(2):
        m.forward_boards(fb, precision="fp16x2")
    torch.cuda.synchronize()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(3):
        m.forward_boards(fb, precision="fp16x2")
    p1.record()
    torch.cuda.synchronize()
    blob2 = m.blob("fp16x2")
    org1 = origins[rank * args.nn_origins:(rank + 1) * args.nn_origins].contiguous()  # this rank's own origins
    eng.mcts_nn_lines(blob2, "fp16x2", org1, number, 400, 0)
    torch.cuda.synchronize()
    barrier()
    q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    q0.record()
    rp = eng.mcts_nn_lines(blob2, "fp16x2", org1, number, 401, rank * args.nn_origins)
    q1.record()
    torch.cuda.synchronize()
    # whole job in the precise mode: the nn_policy loop above (same origins, sharding, exchange and pipelining) with the
    # precise weight image, `steps` evaluations; moves summed over ranks / slowest rank's device time
    pj_ms, pj_moves, _ = timed_evaluations(blob2, "fp16x2", args.steps, 2, 600)
    d.sum_over_ranks(pj_moves)
    pj_t = torch.tensor([pj_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        torch.distributed.all_reduce(pj_t, op=torch.distributed.ReduceOp.MAX)
    precise_job = float(pj_moves.item()) / (float(pj_t.item()) * 1e-3)
    # config 3 in the precise mode: one main game x 16 mcts_nn(number=50) main moves, device-resident
    eng.selfplay_nn_games(blob2, "fp16x2", 1, number, 500, max_steps=4)
    torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    rg = eng.selfplay_nn_games(blob2, "fp16x2", 1, number, 501, max_steps=16)
    g1.record()
    torch.cuda.synchronize()
    precise = {"forward_boards_per_s_per_gpu": 3 * nb / (p0.elapsed_time(p1) * 1e-3),
               "nn_policy_moves_per_s_per_gpu": float(rp["total_moves"].item()) / (q0.elapsed_time(q1) * 1e-3),
               "nn_policy_moves_per_s_all_gpus": precise_job,  # the nn_policy loop (`steps` pipelined evaluations) in the precise mode
               "mcts_nn_game_ms_per_main_move": g0.elapsed_time(g1) / max(1, int(rg["steps"].sum().item())),
               "note": "precision=\"fp16x2\" (B2048_PREC_FP16X2_TC: fp16 hi+lo weights on L1-L3 and activations on L1-L2, 32 most important input channels): max |dlogp| 0.0077 < 1e-2 on all 150,371 fixture boards "
                       "(tests/test_network_gpu.py); same workloads as forward_only / one root evaluation of this block / mcts_nn_game (one main game, number=%d), one GPU" % number}
    import copy
    torch_base = torch_forward_baselines(pkg, copy.deepcopy(m), fb) if rank == 0 else None
    # config 3 (one main game, mcts_nn number=50 per main move) and config 4 (256 main games in total, sharded)
    log("nn_policy: config 3 / config 4 main games")
    game1 = main_games_block(pkg, blob, 1, number, args.game_steps, args.steps, args.warmup, 5000, 0, 1, barrier)
    sp256 = main_games_block(pkg, blob, args.sp_games, number, args.sp_steps, args.steps, args.warmup, 7000, rank, world, barrier)
    return {"ms": ms, "e2e_ms": t0.elapsed_time(t1), "moves": float(moves.item()), "e2e_moves": float(e2e_moves.item()),
            "fwd_boards_per_s": fwd_boards_per_s, "torch_base": torch_base, "launches": launches, "precise": precise,
            "h2d": G * 8, "d2h": G * 4 * 8, "weights": weights, "game1": game1, "sp256": sp256,
            "workload": f"mcts_nn: {G} origins (fresh boards) x 4 root moves x {number} c64b3 policy lines per step over "
                        f"{world} GPU(s): {how}, [G,4] MIN allreduce per step"}


def train_step_block(pkg, args, rank, world, barrier):
    """trainer.py's optimisation step (batch 2048 per GPU, fp32) through TrainStep.step, the call
    the trainer makes; N > 1 = data parallel with a gradient allreduce (NCCL) between backward and Adam."""
    import numpy as np
    import torch

    B = 2048
    torch.manual_seed(1234)
    m = pkg.network.ConvNet(channels=64, blocks=3).cuda().train()
    st = pkg.trainer.TrainStep(m, lr=1e-3, max_batch=B)
    # synthetic records: random tile exponents 0..7 per cell, random soft targets
    g = torch.Generator(device="cuda").manual_seed(5 + rank)
    boards = torch.randint(0, 2 ** 62, (B * 8,), generator=g, device="cuda", dtype=torch.int64) & 0x7777777777777777
    targets = torch.softmax(torch.randn(B * 8, 4, generator=g, device="cuda"), 1)
    steps = max(10, 4 * args.steps)

    def run(stepper):
        for k in range(4):
            stepper.step(boards[(k % 8) * B:(k % 8 + 1) * B], targets[(k % 8) * B:(k % 8 + 1) * B])
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(steps):
            stepper.step(boards[(k % 8) * B:(k % 8 + 1) * B], targets[(k % 8) * B:(k % 8 + 1) * B])
        e1.record()
        barrier()
        return e0.elapsed_time(e1)

    ms = run(st)
    # launch accounting: kernels of one forward+backward enqueued eagerly (the timed steps replay them as a CUDA graph) + Adam
    lc = pkg._lib.launch_count
    l0 = lc()
    st._enqueue(boards[:B], targets[:B], B)
    per_step_launches = lc() - l0 + 1
    torch.manual_seed(1234)
    mixed = pkg.trainer.TrainStep(pkg.network.ConvNet(channels=64, blocks=3).cuda().train(), lr=1e-3, max_batch=B, precision="mixed")
    ms_mixed = run(mixed)
    return {"ms": ms, "ms_mixed": ms_mixed, "steps": steps, "batch": B, "loss": float(st.loss.item()),
            "launches_per_step": per_step_launches}


def workload_name(games):
    return f"play_fixed_batch: {games} concurrent fixed-order (L,U,D,R) games per GPU played to termination"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--games", type=int, default=1 << 24, help="games per GPU per step (BASELINE config[1]: up to 16M)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nn-origins", type=int, default=256, help="mcts_nn origins per GPU per step (config 4: 256)")
    ap.add_argument("--nn-lines", type=int, default=50, help="policy lines per root move (config 3/4: 50)")
    ap.add_argument("--game-steps", type=int, default=32, help="config 3: main moves of the single mcts_nn main game per step")
    ap.add_argument("--sp-games", type=int, default=256, help="config 4: main games in TOTAL (sharded over the GPUs)")
    ap.add_argument("--sp-steps", type=int, default=8, help="config 4: main moves per main game per step")
    ap.add_argument("--no-nn", action="store_true", help="skip the NN-policy block")
    ap.add_argument("--nn-one-stream", action="store_true", help="nn_policy block: launch the evaluations on one stream (A/B)")
    ap.add_argument("--nn-lanes", type=int, default=2, help="nn_policy block: streams the independent evaluations alternate between")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step block")
    ap.add_argument("--cpu-seconds", type=float, default=4.0, help="wall target of the cpu_baseline sample")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import numpy as np
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the first communicator comes up: point fd 1 at stderr until
        # then, so that stdout carries exactly the one JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    pkg = importlib.import_module("2048nn_b200")
    eng = pkg.engine
    n = args.games
    gid_base = rank * n  # ranks play disjoint games of the same seed

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # synthetic start boards: fresh 2-tile boards (generate_init_tiles distribution), resident in HBM
    start = eng.init_boards(n, 777, gid_base)
    out = (torch.empty(n, dtype=torch.int64, device="cuda"), torch.empty(n, dtype=torch.int32, device="cuda"),
           torch.empty(n, dtype=torch.int32, device="cuda"))
    for w in range(args.warmup):
        eng.playout_fixed(n, 1000 + w, gid_base, start=start, out=out)
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- device-resident timed region: exactly K steps -------------------------------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    moves = torch.zeros((), dtype=torch.int64, device="cuda")
    barrier()
    launches0 = pkg._lib.launch_count()
    e0.record()
    for k in range(args.steps):
        eng.playout_fixed(n, 2000 + k, gid_base, start=start, out=out)
    e1.record()
    barrier()
    main_launches = pkg._lib.launch_count() - launches0  # kernels launched through the C ABI inside the timed region
    ms = e0.elapsed_time(e1)
    # move counts of the timed steps (recomputed outside the timed region from the last step,
    # plus an exact recount per seed would cost K more passes; moves/step is seed-stable to 1e-4)
    step_moves = []
    for k in range(args.steps):
        eng.playout_fixed(n, 2000 + k, gid_base, start=start, out=out)
        step_moves.append(int(out[2].sum().item()))
    total_moves = float(sum(step_moves))

    # ---- end to end through the host-buffer API: H2D of start boards + D2H of results ------
    hp = eng.HostPlayout(n)
    h_start = start.cpu().pin_memory()
    for w in range(max(3, args.warmup)):  # whole steps, metric read included (its first use loads a reduction kernel: 1-70 ms)
        hp(h_start, 3000 + w, gid_base)
        _ = int(hp.d_out[2].sum().item())
    # the box's PCIe rates for these very buffers (reported next to e2e: a slow host link shows up here, not in the kernel)
    pcie = {}
    for name, (dst, src) in {"h2d_gbs": (hp.d_start, h_start), "d2h_gbs": (hp.h_out[1], hp.d_out[1])}.items():
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        dst.copy_(src, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        pcie[name] = src.numel() * src.element_size() / (c0.elapsed_time(c1) * 1e-3) / 1e9
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    e2e_moves = 0
    host_t = []
    for k in range(args.steps):
        h0 = time.perf_counter()
        _, hs, hm = hp(h_start, 4000 + k, gid_base)  # host arrays: score + move count of every game
        h1 = time.perf_counter()
        e2e_moves += int(hp.d_out[2].sum().item()) + 0 * int(hm[-1])  # the step's metric, read on the host
        host_t.append((h1 - h0, time.perf_counter() - h1))
    t1.record()
    barrier()
    e2e_ms = t0.elapsed_time(t1)
    if os.environ.get("B2048_E2E_TRACE"):
        mine = "rank %d: e2e %.1f ms/step, pcie %s, host ms per step (HostPlayout call + metric read): %s" % (
            rank, e2e_ms / args.steps, {k: round(v, 1) for k, v in pcie.items()}, " ".join("%.1f+%.1f" % (a * 1e3, b * 1e3) for a, b in host_t))
        every = [mine]
        if world > 1:
            every = [None] * world
            dist.all_gather_object(every, mine)
        for ln in every:
            log(ln)
    if getattr(hp, "last_events", None):  # B2048_E2E_TRACE=1: per-chunk event times of the last step (diagnostics)
        base = hp.last_events[0][0]
        log("e2e trace (ms from the step's first copy): " + " ".join(
            "c%d[in %.1f-%.1f run %.1f-%.1f out -%.1f]" % ((i,) + tuple(base.elapsed_time(x) for x in ev)) for i, ev in enumerate(hp.last_events)))

    # ---- second headline: c64b3 NN-policy playouts (mcts_nn root evaluations, config 3/4 shape) ----
    log("fixed-order block done")
    # measured denominators of the K2 roofline: what this GPU sustains on LOP3 alone (the ALU pipe) and on LOP3 + IMAD (issue limit)
    pipe_rates = {"alu_lop3": eng.pipe_rate(0), "issue_lop3_imad": eng.pipe_rate(1)} if rank == 0 else None
    nn = nn_policy_block(pkg, args, rank, world, barrier) if not args.no_nn else None
    log("training block")
    tr = train_step_block(pkg, args, rank, world, barrier) if not args.no_train else None
    log("blocks done")
    clocks = sampler.stop() if rank == 0 else None

    sp = nn["sp256"] if nn else None
    tmax = torch.tensor([ms, e2e_ms] + ([nn["ms"], nn["e2e_ms"]] if nn else [0.0, 0.0]) + ([tr["ms"], tr["ms_mixed"]] if tr else [0.0, 0.0])
                        + ([sp["ms"], sp["e2e_ms"]] if sp else [0.0, 0.0]), dtype=torch.float64, device="cuda")
    tot = torch.tensor([total_moves, float(e2e_moves)] + ([sp["moves"], sp["mains"], sp["e2e_moves"], float(sp["launches"]),
                                                           float(sp["h2d"]), float(sp["d2h"])] if sp else [0.0] * 6),
                       dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    tl = tmax.tolist()
    ms, e2e_ms = tl[0], tl[1]
    if nn:
        nn["ms"], nn["e2e_ms"] = tl[2], tl[3]
    if tr:
        tr["ms"], tr["ms_mixed"] = tl[4], tl[5]
    if sp:
        sp["ms"], sp["e2e_ms"] = tl[6], tl[7]
        sp["moves"], sp["mains"], sp["e2e_moves"], sp["launches"], sp["h2d"], sp["d2h"] = tot.tolist()[2:8]
    total_moves, e2e_moves = tot.tolist()[:2]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_kind = measured_peaks()
    value = total_moves / (ms * 1e-3)
    e2e_value = e2e_moves / (e2e_ms * 1e-3)
    sm_count = pkg._lib.load_library().b2048_num_sms(pkg._lib.ctx())
    # roofline of the dominant kernel (playout_fixed_kernel): SM instruction-issue bound, not
    # HBM and not tensor.  achieved = algorithmic thread-instr/s on ONE GPU.
    per_gpu_moves_per_s = value / world
    issue_peak = sm_count * 4 * 32 * peaks["sm_max_mhz"] * 1e6  # thread-instr/s at max SM clock
    achieved = per_gpu_moves_per_s * INSTR_PER_MOVE
    hbm_bytes = 24.0 * n * args.steps  # 8 B start + 16 B results per game
    traffic = None  # DRAM bytes per launch of the dominant kernel from the committed ncu capture (same size only)
    tpath = os.path.join(ROOT, "profiles", "r02_k2_traffic_16m.json")
    measured_instr, executed = None, None
    if os.path.exists(tpath):
        tinfo = json.load(open(tpath))
        measured_instr = tinfo.get("thread_instr_per_move")
        if tinfo.get("games") == n:
            traffic = tinfo["traffic_bytes_per_launch"]
        if measured_instr:
            # the executed view: what the kernel issues, against the issue slots and against the ALU pipe that binds it
            executed = {"kernel": tinfo.get("kernel"), "thread_instr_per_move": measured_instr,
                        "issue_frac_at_max_clock": per_gpu_moves_per_s * measured_instr / issue_peak,
                        # the binding pipe: ALU-pipe lane-slots per move (ncu) x live moves/s vs 16 lanes x 4 x SMs x max clock
                        "alu_pipe_frac_at_max_clock": (per_gpu_moves_per_s * tinfo["alu_pipe_lane_slots_per_move"] / (issue_peak / 2)
                                                       if tinfo.get("alu_pipe_lane_slots_per_move") else None),
                        "ncu_alu_pipe_active_pct": tinfo.get("alu_pipe_active_pct"),
                        "ncu_issue_active_pct": tinfo.get("issue_active_pct"),
                        "ncu_shared_wavefronts_pct_of_peak": tinfo.get("shared_wavefronts_pct_of_peak"),
                        "source": "committed ncu capture (profiles/r02_k2_traffic_16m.json), not measured in this run"}
            if pipe_rates:
                # live probes (b2048_probe_pipe_rate, csrc/probe.cuh): full warps of LOP3 / LOP3+IMAD on every SM, device timed
                executed["measured_pipe_rates"] = {
                    "alu_pipe_thread_instr_per_s": pipe_rates["alu_lop3"], "issue_thread_instr_per_s": pipe_rates["issue_lop3_imad"],
                    "alu_formula_at_max_clock": issue_peak / 2, "issue_formula_at_max_clock": issue_peak,
                    "alu_pipe_frac_of_measured": (per_gpu_moves_per_s * tinfo["alu_pipe_lane_slots_per_move"] / pipe_rates["alu_lop3"]
                                                  if tinfo.get("alu_pipe_lane_slots_per_move") else None),
                    "issue_frac_of_measured": per_gpu_moves_per_s * measured_instr / pipe_rates["issue_lop3_imad"],
                    "note": "probe kernels measured in this run; per-move instruction counts from the committed ncu capture"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(n), "games_per_gpu": n, "moves_per_step_per_gpu": step_moves[0],
                   "order": "LUDR", "l2": "inputs+outputs (24 B/game = %d MB) larger than L2" % (24 * n >> 20),
                   "parallelism": f"games sharded over {world} GPU(s), no data-path collective"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": hp.h2d_bytes, "d2h_bytes_per_step": hp.d2h_bytes,
                "ms_per_step": e2e_ms / args.steps, "pcie_probe": pcie},
        "gpu_launches": main_launches,
        "roofline": {"bound": "sm_issue", "achieved": achieved / 1e12, "peak": issue_peak / 1e12, "unit": "Tinstr/s",
                     "frac": achieved / issue_peak, "traffic": traffic,
                     "traffic_source": "committed ncu capture profiles/r02_k2_traffic_16m.json (same size), not measured in this run",
                     "measured_thread_instr_per_move": measured_instr,
                     "executed": executed,
                     "note": "frac = algorithmic 250 thread-instr/move (SURVEY 8d: the reference-shaped attempt cascade) x moves/s vs %d SM x 4 x 32 x "
                             "%.0f MHz (%s clocks); it exceeds 1 because the pair-table kernel does less work per move than that model "
                             "(`executed`: thread-instr/move and pipe utilisation from the committed ncu capture; the ALU pipe and the "
                             "shared-memory wavefronts, both ~89 %% busy, bind it); HBM sees only %.1f GB/s of %.0f"
                             % (sm_count, peaks["sm_max_mhz"], peak_kind, hbm_bytes / (ms * 1e-3) / 1e9 / world,
                                peaks["hbm_gbs"])},
    }
    if nn:
        nn_value = nn["moves"] / (nn["ms"] * 1e-3)  # moves are already summed over ranks
        tf = nn_value / world * 5128704.0 / 1e12     # valid-tap FLOP per evaluation, one GPU
        peak_tf = peaks["bf16_tflops_sustained"]
        nn_traffic = None  # DRAM bytes per launch from the committed ncu capture of this workload (1 GPU, default sizes)
        npath = os.path.join(ROOT, "profiles", "r02_k4_traffic_nn_block.json")
        if os.path.exists(npath) and args.nn_origins == 256 and args.nn_lines == 50:
            nn_traffic = json.load(open(npath))["traffic_bytes_per_launch"]
        fwd_tf = nn["fwd_boards_per_s"] * 5128704.0 / 1e12
        forward_only = {"boards_per_s_per_gpu": nn["fwd_boards_per_s"], "boards": 1 << 20, "tflops_valid_taps": fwd_tf,
                        "frac_of_sustained_peak": fwd_tf / peaks["bf16_tflops_sustained"]}
        if nn["torch_base"]:
            tb = nn["torch_base"]
            best_torch = max(tb["eager_fp32"], tb["channels_last_bf16_autocast"])
            forward_only["torch_baseline"] = {
                "eager_fp32_boards_per_s": tb["eager_fp32"], "channels_last_bf16_autocast_boards_per_s": tb["channels_last_bf16_autocast"],
                "speedup_vs_stronger_arm": nn["fwd_boards_per_s"] / best_torch, "note": tb["note"]}
        line["nn_policy"] = {
            "metric": NN_METRIC, "value": nn_value, "unit": UNIT,
            "ms_per_step": nn["ms"] / args.steps, "dtype": "fp16 operands / fp32 accumulate (tcgen05)",
            "config": {"workload": nn["workload"], "weights": nn["weights"], "scaling": "weak",
                       "l2": "compute-bound kernel: per-step inputs are 8 B per origin, the 0.5 MB weight image is L2-resident by design"},
            "e2e": {"value": nn["e2e_moves"] / (nn["e2e_ms"] * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": nn["h2d"], "d2h_bytes_per_step": nn["d2h"]},
            "gpu_launches": nn["launches"],
            "forward_only": forward_only,
            "precise_mode": nn["precise"],
            "roofline": {"bound": "tensor", "achieved": tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": tf / peak_tf,
                         "traffic": nn_traffic,
                         "traffic_source": "committed ncu capture profiles/r02_k4_traffic_nn_block.json (default sizes), not measured in this run",
                         "note": "5,128,704 valid-tap FLOP per policy evaluation x evals/s on one GPU vs the %s "
                                 "sustained dense bf16/fp16 GEMM peak; the kernel executes 1.25x that (y-padding rows, K-doubled L0)"
                                 % peak_kind},
        }
        g1, spb = nn["game1"], nn["sp256"]
        line["mcts_nn_game"] = {
            "metric": "BASELINE config 3: one main game driven by mcts_nn(number=%d), device-resident (b2048_selfplay_nn)" % args.nn_lines,
            "ms_per_main_move": g1["ms"] / max(1.0, g1["mains"]), "value": g1["moves"] / (g1["ms"] * 1e-3), "unit": UNIT,
            "main_moves_per_step": g1["mains"] / args.steps, "policy_moves_per_main_move": g1["moves"] / max(1.0, g1["mains"]),
            "e2e": {"value": g1["e2e_moves"] / (g1["e2e_ms"] * 1e-3), "unit": UNIT, "h2d_bytes_per_step": g1["h2d"],
                    "d2h_bytes_per_step": g1["d2h"]},
            "gpu_launches": g1["launches"],
            "config": {"workload": "1 fresh main game x %d main moves per step, each main move = 4 root moves x %d policy lines "
                                   "with the min-move early stop; %s" % (args.game_steps, args.nn_lines, nn["weights"]),
                       "parallelism": "replicas only: one main game is 200 lines; with three-ply speculation (each line's board, its "
                                      "four successors and up to 11 of their successors in one forward: up to three policy moves per "
                                      "forward) they occupy 100 of 148 SMs and the call is bound by the longest line's dependent "
                                      "forwards (%d GPU(s) run the same game)" % world}}
        line["selfplay256"] = {
            "metric": "BASELINE config 4: %d main games in TOTAL x mcts_nn(number=%d), device-resident, sharded by game" % (args.sp_games, args.nn_lines),
            "value": spb["moves"] / (spb["ms"] * 1e-3), "unit": UNIT, "scaling": "strong",
            "ms_per_step": spb["ms"] / args.steps, "main_moves_per_s": spb["mains"] / (spb["ms"] * 1e-3),
            "e2e": {"value": spb["e2e_moves"] / (spb["e2e_ms"] * 1e-3), "unit": UNIT, "h2d_bytes_per_step": spb["h2d"],
                    "d2h_bytes_per_step": spb["d2h"]},
            "gpu_launches": int(spb["launches"]),
            "roofline": {"bound": "tensor", "achieved": spb["moves"] / (spb["ms"] * 1e-3) / world * 5128704.0 / 1e12, "peak": peak_tf,
                         "unit": "TFLOP/s", "frac": spb["moves"] / (spb["ms"] * 1e-3) / world * 5128704.0 / 1e12 / peak_tf, "traffic": None},
            "config": {"workload": "%d main games (fresh boards) x %d main moves per step, %d policy lines per legal root move; main "
                                   "games sharded over %d GPU(s) by game id, no collective on the data path (records gathered at the end)"
                                   % (args.sp_games, args.sp_steps, args.nn_lines, world), "weights": nn["weights"]}}
        if world == 1:
            cores_nn = os.cpu_count() or 1
            log("cpu baseline of the NN path (%d threads)" % cores_nn)
            nv, ms_call, nsample = cpu_nn_rate(args.cpu_seconds, cores_nn)
            line["nn_policy"]["cpu_baseline"] = {"value": nv, "unit": UNIT, "cores": cores_nn, "kind": "port",
                                                 "ms_per_mcts_nn_call": ms_call, "sample": nsample}
    if tr:
        us = tr["ms"] * 1e3 / tr["steps"]
        flop = 2.0 * 7577600.0 * tr["batch"]  # in-bounds conv MACs of forward + data gradient + weight gradient per sample
        fma_peak = sm_count * 128 * 2 * peaks["sm_max_mhz"] * 1e6 / 1e12  # packed FFMA2: 128 fp32 FMA/clk/SM
        line["train_step"] = {
            "metric": "c64b3 training samples/sec (trainer.py step: train-mode forward, KLDiv, backward, Adam)",
            "value": tr["batch"] * world / (us * 1e-6), "unit": "samples/s", "us_per_step": us, "dtype": "fp32",
            "config": {"workload": "batch %d per GPU, fp32, hand-written conv fwd/dgrad/wgrad + BatchNorm batch statistics + "
                                   "Adam; %s" % (tr["batch"], "gradient allreduce over %d GPUs" % world if world > 1 else "single GPU"),
                       "scaling": "weak"},
            "gpu_launches": tr["launches_per_step"] * tr["steps"],
            "mixed_precision": {"us_per_step": tr["ms_mixed"] * 1e3 / tr["steps"],
                                "value": tr["batch"] * world / (tr["ms_mixed"] * 1e-3 / tr["steps"]), "unit": "samples/s",
                                "note": "precision=\"mixed\": forward / data-gradient convolutions on tcgen05 (fp16 / bf16 operands, "
                                        "fp32 accumulate), weight gradient + BatchNorm + Adam fp32"},
            "roofline": {"bound": "fp32_fma", "achieved": flop / (us * 1e-6) / 1e12, "peak": fma_peak, "unit": "TFLOP/s",
                         "frac": flop / (us * 1e-6) / 1e12 / fma_peak, "traffic": None,
                         "note": "15.2 MFLOP of in-bounds convolution work per sample vs 148 SM x 128 FMA/clk (FFMA2) at max clock; "
                                 "the reference's step is PyTorch eager: 3.5 ms fp32 / 2.2 ms TF32 on the same GPU (tools/train_bench.py)"},
        }
    # BASELINE config[0]: one full play_mcts_fixed(number=10) main game from a fresh board (device-resident kernel)
    import time as _time
    pkg.mcts.play_mcts_fixed(number=10, seed=1)
    torch.cuda.synchronize()
    t0w = _time.perf_counter()
    games = [pkg.mcts.play_mcts_fixed(number=10, seed=sd) for sd in (12345, 2, 3)]
    torch.cuda.synchronize()
    dtw = _time.perf_counter() - t0w
    line["mcts_fixed_game"] = {
        "workload": "play_mcts_fixed(number=10, mode=0): 3 full main games from fresh boards, one kernel launch each, wall clock",
        "ms_per_game": dtw / 3 * 1e3, "main_moves": [g[2] for g in games], "scores": [g[1] for g in games],
        "us_per_main_move": dtw * 1e6 / sum(g[2] for g in games),
        "reference": "9.56 s per game for the Python reference (BASELINE.md, survey container CPU)"}
    cores = os.cpu_count() or 1
    if world == 1:
        v, sample = cpu_port_rate(args.cpu_seconds, cores)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
