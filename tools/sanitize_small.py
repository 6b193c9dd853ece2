"""Small invocations of the kernels added around the hot path, for compute-sanitizer (memcheck / racecheck)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("2048nn_b200")
eng = pkg.engine
ng = np.load("tests/golden/c64b3_golden.npz")
sd = {k[len("sd_random/"):]: torch.from_numpy(np.array(ng[k])) for k in ng.files if k.startswith("sd_random/")}
boards = torch.from_numpy(ng["lg_boards"][:40].view(np.int64).copy()).cuda()
y = torch.softmax(torch.randn(40, 4, device="cuda"), 1)
for prec in ("fp32", "mixed"):
    m = pkg.network.ConvNet(64, 3); m.load_state_dict(sd); m.cuda().train()
    st = pkg.trainer.TrainStep(m, lr=1e-3, max_batch=64, precision=prec, use_graph=False)
    for _ in range(2):
        st.step(boards, y)
    torch.cuda.synchronize()
    print(prec, "loss", float(st.loss.item()), flush=True)
r = eng.selfplay_fixed_games(3, 3, 5, times=2, mode=2, max_steps=48)
r1 = eng.selfplay_fixed_games(2, 4, 5, times=1, mode=1, max_steps=48)
r0 = eng.selfplay_fixed_games(2, 4, 5, times=1, mode=0, max_steps=48)
torch.cuda.synchronize()
print("selfplay steps", r["steps"].tolist(), r1["steps"].tolist(), r0["steps"].tolist(), flush=True)
m = pkg.network.ConvNet(64, 3, precision="fp16"); m.load_state_dict(sd); m.cuda().eval()
blobs = torch.stack([m.blob("fp16"), m.blob("fp16")]).contiguous()
pop = eng.playout_nn_population(blobs, "fp16", 5, 3)
acc, mt = eng.game_summary(pop["final"].reshape(-1), pop["score"].reshape(-1), pop["moves"].reshape(-1), want_maxtile=True)
torch.cuda.synchronize()
print("population moves", pop["moves"].tolist(), "summary", acc.tolist()[16:], flush=True)
