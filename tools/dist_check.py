"""torchrun --nproc-per-node N tools/dist_check.py: sharded evaluations over N GPUs (NCCL) equal the single-GPU ones."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
os.environ["NCCL_DEBUG"] = "WARN"
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pkg = importlib.import_module("2048nn_b200")
d, eng = pkg.dist, pkg.engine
ng = np.load("tests/golden/c64b3_golden.npz")
m = pkg.network.ConvNet(64, 3, precision="fp16")
m.load_state_dict({k[len("sd_shipped/"):]: torch.from_numpy(np.array(ng[k])) for k in ng.files if k.startswith("sd_shipped/")})
m.cuda().eval()
blob = m.blob("fp16")
ok = True
for G, number in ((1, 12), (1, 1300), (world, 12), (2 * world + 1, 12)):  # replicated, line-sharded, origin-sharded (even / ragged)
    origins = eng.init_boards(G, 31, 0)
    r = d.mcts_nn_sharded(blob, "fp16", origins, number, 5, 100)
    one = eng.mcts_nn_lines(blob, "fp16", origins, number, 5, 100)
    same = torch.equal(r["min"], d.min_from_stats(one["legal"], one["minmov"]))
    f = d.mcts_fixed_sharded(origins, number, 5, 7)
    f1 = eng.mcts_fixed_lines(origins, number, 5, 7, per_line=False, fused=True)
    same_f = torch.equal(f["logsum"], f1["logsum"]) and torch.equal(f["count"], f1["count"]) and torch.equal(f["minmov"], f1["minmov"])
    ok &= same and same_f
    if rank == 0:
        print(f"G={G} number={number}: plan {d.shard_plan(G, number, rank, world)} mcts_nn {'ok' if same else 'MISMATCH'}, mcts_fixed {'ok' if same_f else 'MISMATCH'}", flush=True)
# pipelined root evaluations: three [G,4] exchanges in flight (async_op) before the first is waited for
origins = eng.init_boards(2 * world + 1, 31, 0)
want = [d.mcts_nn_sharded(blob, "fp16", origins, 12, 50 + k, 100 * k)["min"].clone() for k in range(3)]
pend = [d.mcts_nn_sharded(blob, "fp16", origins, 12, 50 + k, 100 * k, async_reduce=True) for k in range(3)]
same_p = all(torch.equal(d.mcts_nn_finish(r)["min"], w) for r, w in zip(pend, want))
ok &= same_p
if rank == 0:
    print(f"pipelined mcts_nn (3 exchanges in flight) {'ok' if same_p else 'MISMATCH'}", flush=True)
recs, mv = pkg.selfplay.selfplay_batch(2 * world + 1, 6, None, 2, seed=9)
single = eng.selfplay_fixed_games(2 * world + 1, 6, 9, times=2, mode=2)
steps = single["steps"].cpu().numpy()
hb = single["boards"].cpu().numpy().view(np.uint64)
same_sp = all(np.array_equal(recs[g]["boards"], hb[g, :steps[g]]) for g in range(2 * world + 1)) and mv == int(single["total_moves"].item())
ok &= same_sp
if rank == 0:
    print(f"selfplay_batch sharded over {world} ranks {'ok' if same_sp else 'MISMATCH'} ({mv} playout moves)", flush=True)
# NN-policy main games (config 4 shape): device-resident, sharded by game over the ranks == one GPU playing them all
mr = pkg.network.ConvNet(64, 3, precision="fp16")
mr.load_state_dict({k[len("sd_random/"):]: torch.from_numpy(np.array(ng[k])) for k in ng.files if k.startswith("sd_random/")})
mr.cuda().eval()
Gn = 2 * world + 1
recs_nn, mv_nn = pkg.selfplay.selfplay_batch(Gn, 5, mr, seed=21, precision="fp16")
one_nn = eng.selfplay_nn_games(mr.blob("fp16"), "fp16", Gn, 5, 21, max_steps=16384)
st_nn = one_nn["steps"].cpu().numpy()
hb_nn = one_nn["boards"].cpu().numpy().view(np.uint64)
hr_nn = one_nn["results"].cpu().numpy()
same_nn = all(np.array_equal(recs_nn[g]["boards"], hb_nn[g, :st_nn[g]]) and np.array_equal(recs_nn[g]["results"], hr_nn[g, :st_nn[g]])
              for g in range(Gn))
ok &= same_nn
if rank == 0:
    print(f"NN-policy selfplay_batch ({Gn} main games, mcts_nn number=5) sharded over {world} ranks {'ok' if same_nn else 'MISMATCH'} "
          f"({mv_nn} policy moves, {int(st_nn.sum())} main moves)", flush=True)
# data-parallel training step: every rank ends with identical parameters
torch.manual_seed(0)
st = pkg.trainer.TrainStep(pkg.network.ConvNet(64, 3).cuda().train(), lr=1e-3, max_batch=256)
g = torch.Generator(device="cuda").manual_seed(100 + rank)
for _ in range(3):
    b = torch.randint(0, 2 ** 62, (256,), generator=g, device="cuda", dtype=torch.int64) & 0x7777777777777777
    st.step(b, torch.softmax(torch.randn(256, 4, generator=g, device="cuda"), 1))
st.sync_buffers()
ref, refb = st.params.clone(), st.bn.clone()
dist.broadcast(ref, 0)
dist.broadcast(refb, 0)
same_tr = torch.equal(ref, st.params) and torch.equal(refb, st.bn)
ok &= same_tr
if rank == 0:
    print(f"data-parallel TrainStep: parameters and BatchNorm buffers identical on all ranks: {same_tr}", flush=True)
flag = torch.tensor([0 if ok else 1], device="cuda")
dist.all_reduce(flag)
dist.destroy_process_group()
sys.exit(1 if int(flag.item()) else 0)
