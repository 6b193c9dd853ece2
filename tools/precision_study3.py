"""CPU emulation of the precise tensor-core mode with correction passes over the 32 most important input channels
(csrc/convnet_tc.cuh SPLIT_NK = 2, 2048nn_b200/convnet.split_channel_order) on the 150k-board fixture.

Calibration boards are built the way network.calibration_boards builds them on the GPU (fixed-order games at depths
0..159), here with the oracle's engine.  Also checks that convnet.permute_channels leaves the network's function unchanged.

  python tools/precision_study3.py [n_boards]
"""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F

from oracle import oracle

cn = importlib.import_module("2048nn_b200.convnet")
torch.set_num_threads(os.cpu_count() or 1)
ng = np.load("tests/golden/c64b3_golden.npz")
fx = np.load("tests/golden/c64b3_logits_100k.npz")
N = int(sys.argv[1]) if len(sys.argv) > 1 else fx["boards"].size
sd = {k[len("sd_shipped/"):]: torch.from_numpy(np.array(ng[k])) for k in ng.files if k.startswith("sd_shipped/")}
folded = cn.fold_state_dict(sd)


def calibration_boards(n=2048, seed=0x2048CA11, max_depth=160):
    keys = np.array([oracle.rng_key(seed, g) for g in range(n)], np.uint64)
    boards = np.array([oracle.init_board(int(k)) for k in keys], np.uint64)
    scores, sidx, alive = np.zeros(n, np.uint64), np.zeros(n, np.uint32), np.ones(n, np.uint8)
    prio = np.tile(np.array([0, 1, 3, 2], np.uint8), (n, 1))
    depth = (np.arange(n) * 7919) % max_depth
    snap = boards.copy()
    for step in range(1, max_depth):
        oracle.policy_step(boards, scores, sidx, keys, prio, alive)
        snap = np.where(depth == step, boards, snap)
    return snap


def r16(x):
    return x.half().float()


def forward_fp32(fd, x):
    w, b = [torch.from_numpy(np.asarray(a, np.float32)) for a in fd["conv_w"]], [torch.from_numpy(np.asarray(a, np.float32)).view(1, -1, 1, 1) for a in fd["conv_b"]]
    x = F.relu(F.conv2d(x, w[0], padding=1) + b[0])
    for blk in range(3):
        h = F.relu(F.conv2d(x, w[1 + 2 * blk], padding=1) + b[1 + 2 * blk])
        x = F.relu(F.conv2d(h, w[2 + 2 * blk], padding=1) + b[2 + 2 * blk] + x)
    hw = torch.from_numpy(np.asarray(fd["head_w"], np.float32)).view(4, 64, 1, 1)
    x = F.relu(F.conv2d(x, hw) + torch.from_numpy(np.asarray(fd["head_b"], np.float32)).view(1, -1, 1, 1))
    z = F.linear(x.reshape(-1, 64), torch.from_numpy(np.asarray(fd["fc_w"], np.float32)), torch.from_numpy(np.asarray(fd["fc_b"], np.float32)))
    return F.log_softmax(z, 1)


def emulate(fd, x, split_layers, nch):
    """kernel arithmetic on the (already permuted) network: hi*hi everywhere, + lo*hi and hi*lo over input channels [0, nch) on split layers"""
    w = [torch.from_numpy(np.asarray(a, np.float32)) for a in fd["conv_w"]]
    b = [torch.from_numpy(np.asarray(a, np.float32)).view(1, -1, 1, 1) for a in fd["conv_b"]]

    def conv(l, xin):
        wh = r16(w[l]); wl = r16(w[l] - wh); xh = r16(xin); xl = r16(xin - xh)
        y = F.conv2d(xh, wh, padding=1)
        if l in split_layers:
            y = y + F.conv2d(xl[:, :nch], wh[:, :nch], padding=1) + F.conv2d(xh[:, :nch], wl[:, :nch], padding=1)
        return y + b[l]

    x = F.relu(F.conv2d(x, w[0], padding=1) + b[0])  # L0 exact
    for blk in range(3):
        l1, l2 = 1 + 2 * blk, 2 + 2 * blk
        h = F.relu(conv(l1, x))
        x = F.relu(conv(l2, h) + r16(x))
        if blk < 2 and (l2 + 1) not in split_layers:
            x = r16(x)
    hw = torch.from_numpy(np.asarray(fd["head_w"], np.float32)).view(4, 64, 1, 1)
    x = F.relu(F.conv2d(x, hw) + torch.from_numpy(np.asarray(fd["head_b"], np.float32)).view(1, -1, 1, 1))
    z = F.linear(x.reshape(-1, 64), torch.from_numpy(np.asarray(fd["fc_w"], np.float32)), torch.from_numpy(np.asarray(fd["fc_b"], np.float32)))
    return F.log_softmax(z, 1)


def emulate2(fd, x, w_layers, a_layers, nch):
    """as emulate(), with separate layer sets for the weight (hi*lo) and activation (lo*hi) correction passes"""
    w = [torch.from_numpy(np.asarray(a, np.float32)) for a in fd["conv_w"]]
    b = [torch.from_numpy(np.asarray(a, np.float32)).view(1, -1, 1, 1) for a in fd["conv_b"]]

    def conv(l, xin):
        wh = r16(w[l]); wl = r16(w[l] - wh); xh = r16(xin); xl = r16(xin - xh)
        y = F.conv2d(xh, wh, padding=1)
        if l in a_layers:
            y = y + F.conv2d(xl[:, :nch], wh[:, :nch], padding=1)
        if l in w_layers:
            y = y + F.conv2d(xh[:, :nch], wl[:, :nch], padding=1)
        return y + b[l]

    x = F.relu(F.conv2d(x, w[0], padding=1) + b[0])
    for blk in range(3):
        l1, l2 = 1 + 2 * blk, 2 + 2 * blk
        h = F.relu(conv(l1, x))
        x = F.relu(conv(l2, h) + r16(x))
        if blk < 2 and (l2 + 1) not in a_layers:
            x = r16(x)
    hw = torch.from_numpy(np.asarray(fd["head_w"], np.float32)).view(4, 64, 1, 1)
    x = F.relu(F.conv2d(x, hw) + torch.from_numpy(np.asarray(fd["head_b"], np.float32)).view(1, -1, 1, 1))
    z = F.linear(x.reshape(-1, 64), torch.from_numpy(np.asarray(fd["fc_w"], np.float32)), torch.from_numpy(np.asarray(fd["fc_b"], np.float32)))
    return F.log_softmax(z, 1)


def report(name, lp, ref):
    d = (lp - ref).abs().max(1).values
    agree = (lp.argmax(1) == ref.argmax(1)).float().mean().item()
    print(f"  {name:60s} max={d.max():.5f} p99.99={d.quantile(0.9999):.5f} >1e-2: {(d >= 1e-2).sum().item():4d} "
          f">5e-3: {(d >= 5e-3).sum().item():5d} argmax={100 * agree:.4f}%", flush=True)


cal = calibration_boards()
stream, mid = cn.split_channel_order(folded, cal)
perm = cn.permute_channels(folded, stream, mid)
x = torch.from_numpy(cn.onehot_np(fx["boards"][:N]))
ref = torch.from_numpy(fx["lg_shipped"][:N])
with torch.no_grad():
    a, b = forward_fp32(folded, x[:4000]), forward_fp32(perm, x[:4000])
    print("permuted network == network (fp32, 4000 boards): max |d| %.2e; vs the reference's log-probs %.2e" % ((a - b).abs().max(), (a - ref[:4000]).abs().max()))
    print("stream order (first 32):", stream[:32].tolist())
    report("fp16, no correction passes (throughput mode)", emulate(perm, x, (), 0), ref)
    report("L1-3 hi+lo, all 64 channels (round-2 first version, 13 passes)", emulate(perm, x, (1, 2, 3), 64), ref)
    report("L1-3 hi+lo, first 32 channels of the calibrated order (10 passes)", emulate(perm, x, (1, 2, 3), 32), ref)
    report("THE KERNEL: weights L1-3, activations L1-2, first 32 channels (9.5 passes)", emulate2(perm, x, (1, 2, 3), (1, 2), 32), ref)
    report("L1-3 hi+lo, first 16 channels (8.5 passes)", emulate(perm, x, (1, 2, 3), 16), ref)
    report("L1-3 hi+lo, first 32 channels, NO reordering", emulate(folded, x, (1, 2, 3), 32), ref)


if os.environ.get("B2048_STUDY_MORE"):
    with torch.no_grad():
        report("W L1-2 + A L1-3, first 32 channels (9.5 passes)", emulate2(perm, x, (1, 2), (1, 2, 3), 32), ref)
        report("W L1-2 + A L1-2, first 32 channels (9 passes)", emulate2(perm, x, (1, 2), (1, 2), 32), ref)
        report("W L1-3 + A L1-3, first 48 channels (11.5 passes)", emulate2(perm, x, (1, 2, 3), (1, 2, 3), 48), ref)
