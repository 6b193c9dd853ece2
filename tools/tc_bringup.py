"""Bring-up of the tcgen05 forward: per-layer activations and logits vs the fp32 torch oracle."""
import ctypes as C
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F

from oracle import oracle

pkg = importlib.import_module("2048nn_b200")
L = pkg._lib
cn = importlib.import_module("2048nn_b200.convnet")
ng = np.load("tests/golden/c64b3_golden.npz")
tag = sys.argv[1] if len(sys.argv) > 1 else "random"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
sd = {k[len(f"sd_{tag}/"):]: torch.from_numpy(np.array(ng[k])) for k in ng.files if k.startswith(f"sd_{tag}/")}
folded = cn.fold_state_dict(sd)
blob = cn.pack_tc(folded).cuda()
boards_np = ng["lg_boards"][:n]
boards = torch.from_numpy(boards_np.view(np.int64).copy()).cuda()

# reference activations (fp32, BN folded)
x = oracle.onehot(boards_np)
acts = []
with torch.no_grad():
    w = [torch.from_numpy(a).float() for a in folded["conv_w"]]
    b = [torch.from_numpy(a).float() for a in folded["conv_b"]]
    h = F.relu(F.conv2d(x, w[0], padding=1) + b[0].view(1, -1, 1, 1)); acts.append(h)
    for blk in range(3):
        t = F.relu(F.conv2d(h, w[1 + 2 * blk], padding=1) + b[1 + 2 * blk].view(1, -1, 1, 1)); acts.append(t)
        h = F.relu(h + F.conv2d(t, w[2 + 2 * blk], padding=1) + b[2 + 2 * blk].view(1, -1, 1, 1)); acts.append(h)
ref_logp = ng[f"lg_{tag}"][:n]
lib = L.load_library()
for layer in list(range(7)):
    dbg = torch.zeros((n, 64, 16), dtype=torch.float32, device="cuda")
    logits = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
    L.check(lib.b2048_convnet_tc_debug(L.ctx(), L.ptr(blob), L.ptr(boards), L.ptr(logits), L.ptr(dbg), layer, n, L.stream_ptr()))
    torch.cuda.synchronize()
    got = dbg.cpu().reshape(n, 64, 4, 4)
    ref = acts[layer]
    err = (got - ref).abs()
    print(f"layer {layer}: max|ref|={ref.abs().max():.3f} max err={err.max():.5f} mean err={err.mean():.6f} "
          f"nonzero got={int((got != 0).sum())} ref={int((ref != 0).sum())}", flush=True)
lp = torch.log_softmax(logits, 1).cpu().numpy()
d = np.abs(lp - ref_logp)
print(f"logp: max|d|={d.max():.5f} argmax agree={(lp.argmax(1) == ref_logp.argmax(1)).mean() * 100:.2f}%")
m = pkg.network.ConvNet(64, 3)
m.load_state_dict(sd); m.cuda().eval()
lp2 = m.forward_boards(boards, precision="fp16").cpu().numpy()
print("forward_boards fp16: max|d| vs ref", np.abs(lp2 - ref_logp).max())
