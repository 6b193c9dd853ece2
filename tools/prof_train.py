"""Three fused training steps at batch 2048 for an ncu launch list."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("2048nn_b200")
B = 2048
ng = np.load("tests/golden/c64b3_golden.npz")
rng = np.random.default_rng(0)
boards = torch.from_numpy(ng["lg_boards"][rng.integers(0, len(ng["lg_boards"]), B)].view(np.int64)).cuda()
y = torch.softmax(torch.randn(B, 4, device="cuda"), 1)
torch.manual_seed(0)
st = pkg.trainer.TrainStep(pkg.network.ConvNet(64, 3).cuda().train(), lr=1e-3, max_batch=B, use_graph=False,
                            precision=sys.argv[1] if len(sys.argv) > 1 else "fp32")
for _ in range(3):
    st.step(boards, y)
torch.cuda.synchronize()
print("ok", float(st.loss.item()))
