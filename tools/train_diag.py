"""Gradient error of the fused fp32 step and of PyTorch's own fp32 autograd, both against fp64 autograd."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch, torch.nn as nn
import test_trainer_gpu as T
pkg = importlib.import_module("2048nn_b200")
ng = np.load("tests/golden/c64b3_golden.npz")
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ours, ref = T.make_pair(pkg, 1, torch.float64)
_, ref32 = T.make_pair(pkg, 1, torch.float32)
boards, y = T.batch(ng, B, 2)
st = pkg.trainer.TrainStep(ours, max_batch=2048)
loss = st.forward_backward(boards, y)
l64 = nn.KLDivLoss(reduction='batchmean')(T.torch_forward(ref, T.onehot(boards).double()), y.double()); l64.backward()
l32 = nn.KLDivLoss(reduction='batchmean')(T.torch_forward(ref32, T.onehot(boards).float()), y); l32.backward()
print("loss", float(loss), float(l64), float(l32))
off = 0
for (name, p), (_, q) in zip(ref.named_parameters(), ref32.named_parameters()):
    n = p.numel(); mine = st.grad[off:off+n].view(p.shape).double(); off += n
    sc = p.grad.abs().max().item()
    print(f"{name:28s} scale {sc:.3e}  ours {(mine-p.grad).abs().max().item()/sc:.2e}  torch32 {(q.grad.double()-p.grad).abs().max().item()/sc:.2e}")
