"""Self-play record generation (fqjin/2048NN selfplay.py) for a BATCH of concurrent main games.

`selfplay` / `selfplay_fixed` keep the reference's single-game signatures and npz format
(selfplay.py:55-58,111-114: boards uint64[T], results float32[T,4], score).  `selfplay_batch`
is the B200 shape of BASELINE config 4: G main games played DEVICE-RESIDENT -- fixed-order search:
one CTA per main game (b2048_selfplay_fixed); NN-policy search: one persistent tensor-core launch whose
game slots are fed by a device-side queue of root-evaluation lines (b2048_selfplay_nn), so main games advance
independently with no host round trip and no barrier per main move.  Main games are sharded over the process
group's GPUs by game id; the only collectives are the final gather of the records and the move-count sum.
`_selfplay_batch_steps` is the lock-step host loop (one root-evaluation launch + [G,4] allreduce per main step,
lines or origins sharded by dist.shard_plan) that the device-resident paths are tested against."""
import os

import numpy as np
import torch

from . import dist as _dist
from . import engine
from .board import _raise_if_flag, _session_seed
from .network import blob_for


def _eval_ids(step, G):
    # evaluation id of main game g at main step t: disjoint stream ranges per (t, g)
    return (step + 1) * G


def selfplay_batch(G, number=10, model=None, times=1, seed=None, precision=None, starts=None, max_steps=100000,
                   group=None, game_id_base=0, games_total=None, lockstep=False):
    """Play G main games to the end.  model=None -> fixed-order playouts scored by
    mcts_fixed_min averaged over `times` (selfplay.py:9-59); else NN-policy mcts_nn (selfplay.py:62-115).

    Returns a list of G records dict(boards uint64[T], results float32[T,4], score int, truncated bool) and the
    total number of playout moves (all ranks).  `truncated` marks a game that was still running after max_steps
    main moves (its record is a prefix and must not be saved as a finished game).

    Main game k of the call is game id game_id_base + k of a session of games_total games (default: this call's G):
    the same (seed, game id, games_total) always gives the same game, whatever the batch it is played in and however
    the batch is sharded over GPUs.  lockstep=True forces the step-by-step host loop."""
    s = _session_seed(seed)
    if lockstep or (model is None and times > 64):
        if game_id_base or games_total not in (None, G):
            raise ValueError("the lock-step loop plays whole sessions (game_id_base = 0, games_total = G)")
        return _selfplay_batch_steps(G, number, model, times, seed, precision, starts, max_steps, group)
    gt = game_id_base + G if games_total is None else int(games_total)
    if model is None:
        def launch(n_local, g0, st, budget):
            return engine.selfplay_fixed_games(n_local, int(number), s, times=int(times), mode=2, starts=st,
                                               game_id_base=game_id_base + g0, games_total=gt, max_steps=budget)
        first_budget = 4096
    else:
        blob, prec = blob_for(model, precision)

        def launch(n_local, g0, st, budget):
            return engine.selfplay_nn_games(blob, prec, n_local, int(number), s, starts=st, game_id_base=game_id_base + g0,
                                            games_total=gt, max_steps=budget)
        first_budget = 16384  # a strong policy's main games run to several thousand moves; a replay costs minutes
    return _device_games(launch, G, starts, max_steps, first_budget, group)


def _device_games(launch, G, starts, max_steps, first_budget, group):
    """Shard G device-resident main games over the process group, replay with a larger step budget while some
    game outlives it, gather every rank's records."""
    import torch.distributed as td
    dev = torch.device("cuda")
    world = td.get_world_size(group) if td.is_available() and td.is_initialized() else 1
    rank = td.get_rank(group) if world > 1 else 0
    g0, n_local, per = _dist.shard_games(G, rank, world)
    budget = int(min(max_steps, first_budget))
    while True:
        r = None
        if n_local:
            st = None if starts is None else starts.to(dev)[g0:g0 + n_local].contiguous()
            r = launch(n_local, g0, st, budget)
        _raise_if_flag()
        clipped = torch.zeros(1, dtype=torch.int32, device=dev)
        if r is not None:
            clipped += (r["status"] != 0).any().to(torch.int32)
        if world > 1:
            td.all_reduce(clipped, op=td.ReduceOp.MAX, group=group)
        if not int(clipped.item()) or budget >= max_steps:
            break
        budget = int(min(max_steps, budget * 8))  # some game outlived the step budget: replay with a larger one
    # gather the shards (padded to `per` games) so that every rank returns all G records
    def padded(t, shape, dtype):
        out = torch.zeros((per,) + shape, dtype=dtype, device=dev)
        if n_local:
            out[:n_local] = t
        return out
    steps = padded(r["steps"] if r else None, (), torch.int32)
    score = padded(r["score"] if r else None, (), torch.int64)
    status = padded(r["status"] if r else None, (), torch.int32)
    boards = padded(r["boards"] if r else None, (budget,), torch.int64)
    results = padded(r["results"] if r else None, (budget, 4), torch.float32)
    moves = r["total_moves"].clone() if r else torch.zeros(1, dtype=torch.int64, device=dev)
    if world > 1:
        def gather(t):
            out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=dev)
            td.all_gather_into_tensor(out, t.contiguous(), group=group)
            return out.reshape((world * per,) + tuple(t.shape[1:]))
        steps, score, status, boards, results = gather(steps), gather(score), gather(status), gather(boards), gather(results)
        td.all_reduce(moves, group=group)
    hs, hsc, hst = steps.cpu().numpy(), score.cpu().numpy(), status.cpu().numpy()
    hb, hr = boards.cpu().numpy().view(np.uint64), results.cpu().numpy()
    records = [dict(boards=hb[g, :hs[g]].copy(), results=hr[g, :hs[g]].copy(), score=int(hsc[g]), truncated=bool(hst[g]))
               for g in range(G)]
    return records, int(moves.item())


def _selfplay_batch_steps(G, number=10, model=None, times=1, seed=None, precision=None, starts=None, max_steps=100000,
                          group=None):
    """Lock-step main games: every main step is one root-evaluation launch for all G games (lines sharded over the
    ranks, one [G,4] allreduce) plus tiny device ops; the host reads one scalar per main step."""
    s = _session_seed(seed)
    dev = torch.device("cuda")
    boards = engine.init_boards(G, s, 0, device=dev) if starts is None else starts.to(dev).clone()
    gids = torch.arange(G, dtype=torch.int64, device=dev)
    alive = torch.ones(G, dtype=torch.bool, device=dev)
    score = torch.zeros(G, dtype=torch.int64, device=dev)
    kmain = torch.zeros(G, dtype=torch.int32, device=dev)
    hist_b, hist_r, hist_alive = [], [], []
    total_moves = torch.zeros(1, dtype=torch.int64, device=dev)
    blob = prec = None
    if model is not None:
        blob, prec = blob_for(model, precision)
    step = 0
    while step < max_steps:
        if model is None:
            res = torch.zeros((G, 4), dtype=torch.float64, device=dev)
            for t in range(times):
                r = _dist.mcts_fixed_sharded(boards, number, s, _eval_ids(step * times + t, G), group)
                res += r["min"].to(torch.float64)
            res /= times
        else:
            r = _dist.mcts_nn_sharded(blob, prec, boards, number, s, _eval_ids(step, G), group)
            res = r["min"].to(torch.float64)
            total_moves += r["total_moves"]
        choice = res.argmax(dim=1).to(torch.uint8)  # np.argmax: first maximum (selfplay.py:39,95)
        nb, gain, moved = engine.move_batch(boards, choice)
        adv = alive & moved.to(torch.bool)  # game over when the argmax move is illegal (selfplay.py:40-42)
        hist_b.append(boards.clone())
        hist_r.append(res.to(torch.float32))
        hist_alive.append(adv.clone())
        spawned, _ = engine.spawn_batch(torch.where(adv, nb, boards), s, gids=gids, k=kmain)
        boards = torch.where(adv, spawned, boards)
        score += torch.where(adv, gain.to(torch.int64), torch.zeros_like(score))
        kmain += adv.to(torch.int32)
        alive = adv
        step += 1
        if not bool(alive.any().item()):  # the one host read per main step
            break
    _dist.sum_over_ranks(total_moves, group)
    hb = torch.stack(hist_b).cpu().numpy().view(np.uint64)  # [T, G]
    hr = torch.stack(hist_r).cpu().numpy()
    ha = torch.stack(hist_alive).cpu().numpy()
    sc = score.cpu().numpy()
    records = []
    for g in range(G):
        T = int(ha[:, g].sum())  # steps are contiguous from 0 while the game advances
        records.append(dict(boards=hb[:T, g].copy(), results=hr[:T, g].copy(), score=int(sc[g]),
                            truncated=bool(step >= max_steps and ha[-1, g])))
    return records, int(total_moves.item())


def save_record(path, rec):
    """selfplay.py:55-58 / 111-114 npz format.  A record cut off by the step budget is not a finished game."""
    if rec.get("truncated"):
        raise ValueError("refusing to save a truncated self-play record (raise max_steps)")
    np.savez(path, boards=np.asarray(rec["boards"], dtype=np.uint64), results=np.asarray(rec["results"], dtype=np.float32),
             score=rec["score"])


CLI_SESSION_SEED = 2048       # command-line records: stream seed shared by every record ...
CLI_GAMES_TOTAL = 1 << 32     # ... of one fixed session of 2^32 main games; --seed N selects main game N of it


def selfplay_fixed(name, board=None, number=10, times=1, verbose=False, seed=None, out_dir='selfplay/fixed10/',
                   game_id=0, games_total=None):
    """selfplay.py:9-59 (single main game, fixed-order min-move search)."""
    print(f'fixed {number} lines, {times} times')
    starts = None if not board else torch.from_numpy(np.array([board], np.uint64).view(np.int64))
    recs, _ = selfplay_batch(1, number, None, times, seed, starts=starts, game_id_base=game_id, games_total=games_total)
    return _report_and_save(recs[0], out_dir, name, 'Saved as {}.npz')


def selfplay(name, model, board=None, number=10, device='cuda', verbose=False, seed=None, out_dir='selfplay/',
             game_id=0, games_total=None):
    """selfplay.py:62-115 (single main game, mcts_nn search)."""
    print('Seed {}'.format(name))
    starts = None if not board else torch.from_numpy(np.array([board], np.uint64).view(np.int64))
    recs, _ = selfplay_batch(1, number, model, 1, seed, starts=starts, game_id_base=game_id, games_total=games_total)
    return _report_and_save(recs[0], out_dir, name, 'Saved {}.npz')


def _report_and_save(rec, out_dir, name, fmt):
    print('Score {}'.format(rec["score"]))
    print('{} moves'.format(len(rec["boards"])))
    os.makedirs(out_dir, exist_ok=True)
    save_record(os.path.join(out_dir, name), rec)
    print(fmt.format(name))
    return rec


def main(argv=None):
    """Command line of selfplay.py:118-152.  `--seed N` names the record NNNNN.npz as the reference does; here N is
    the main-game id inside one fixed session (stream seed CLI_SESSION_SEED, CLI_GAMES_TOTAL games), so record N is
    the same game whether it is produced alone (`--seed N`) or as part of a batch (`--seed M --games K`, records
    M..M+K-1 in one device-resident launch).  Default search: fixed-order min-move, number=10, times=4
    (the reference's default branch) -> selfplay/fixed10/; `--model PATH` -> mcts_nn search -> selfplay/."""
    from argparse import ArgumentParser
    from time import time

    parser = ArgumentParser()
    parser.add_argument('--seed', type=int, default=12345, help='main-game id = save name of the (first) record')
    parser.add_argument('--verbose', action='store_true', help='Verbose output')
    parser.add_argument('--games', type=int, default=1, help='records to generate: main games seed..seed+games-1')
    parser.add_argument('--model', type=str, default='', help='c64b3 state_dict (.pt): NN-policy search instead of fixed-order')
    parser.add_argument('--number', type=int, default=10)
    args = parser.parse_args(argv)
    if args.seed < 0:
        raise ValueError('Seed is {}'.format(args.seed))
    start_time = time()
    model, out_dir, times = None, 'selfplay/fixed10/', 4
    if args.model:
        from .network import ConvNet
        model = ConvNet(channels=64, blocks=3, precision="fp16")
        model.load_state_dict(torch.load(args.model, map_location="cpu"))
        model.to('cuda').eval()
        out_dir, times = 'selfplay/', 1
        print(f'mcts_nn {args.number} lines, {args.games} games')
    else:
        print(f'fixed {args.number} lines, {times} times, {args.games} games')
    recs, _ = selfplay_batch(args.games, args.number, model, times, seed=CLI_SESSION_SEED, game_id_base=args.seed,
                             games_total=CLI_GAMES_TOTAL)
    os.makedirs(out_dir, exist_ok=True)
    for k, rec in enumerate(recs):
        save_record(os.path.join(out_dir, str(args.seed + k).zfill(5)), rec)
        print('Saved {}.npz: score {}, {} moves'.format(str(args.seed + k).zfill(5), rec["score"], len(rec["boards"])))
    total_time = time() - start_time
    print(f'Took {total_time/60:.1f} minutes')


if __name__ == '__main__':
    main()
