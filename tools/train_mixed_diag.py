"""Mixed-mode (tensor-core convs) training step against fp64 autograd and against our fp32 mode."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch, torch.nn as nn
import test_trainer_gpu as T
pkg = importlib.import_module("2048nn_b200")
ng = np.load("tests/golden/c64b3_golden.npz")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ours, ref = T.make_pair(pkg, 1, torch.float64)
boards, y = T.batch(ng, B, 2)
st = pkg.trainer.TrainStep(ours, max_batch=2048, precision=sys.argv[2] if len(sys.argv) > 2 else "mixed", use_graph=False)
loss = st.forward_backward(boards, y)
torch.cuda.synchronize()
lp = T.torch_forward(ref, T.onehot(boards).double())
l64 = nn.KLDivLoss(reduction='batchmean')(lp, y.double()); l64.backward()
print("loss mixed", float(loss), "fp64", float(l64), "max|dlogp|", float((st.log_probs(B).double() - lp.detach()).abs().max()))
off = 0
for name, p in ref.named_parameters():
    n = p.numel(); mine = st.grad[off:off+n].view(p.shape).double(); off += n
    g = p.grad
    cos = float((mine * g).sum() / (mine.norm() * g.norm() + 1e-30))
    print(f"{name:28s} rel L2 err {float((mine-g).norm()/g.norm()):.3e}  cos {cos:.6f}")
