"""c64b3 training step (trainer.py:55-62) at batch 2048: the hand-written fp32 step vs PyTorch eager (cuDNN) on the
same GPU.  Device-timed with CUDA events."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn as nn
pkg = importlib.import_module("2048nn_b200")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
ng = np.load("tests/golden/c64b3_golden.npz")
rng = np.random.default_rng(0)
boards = torch.from_numpy(ng["lg_boards"][rng.integers(0, len(ng["lg_boards"]), B)].view(np.int64)).cuda()
y = torch.softmax(torch.randn(B, 4, device="cuda"), 1)
def timed(fn, n=50, warm=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
torch.manual_seed(0)
m = pkg.network.ConvNet(64, 3).cuda().train()
st = pkg.trainer.TrainStep(m, lr=1e-3, max_batch=B)
us = timed(lambda: st.step(boards, y))
flop = 3 * 2 * B * (16 * 9 * 16 * 64 + 6 * 16 * 9 * 64 * 64)  # dense (zero-padded) conv MACs x2, fwd + dgrad + wgrad
print(f"b200 fused step  bs={B}: {us:.0f} us/step  ({flop/us/1e6:.1f} TFLOP/s dense-conv equivalent)", flush=True)
# CUDA graph of the whole step
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    st.step(boards, y); torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        L = pkg._lib
        L.check(L.load_library().b2048_train_forward_backward(L.ctx(st.device), L.ptr(st.params), L.ptr(st.bn), L.ptr(st.workspace),
                st.max_batch, L.ptr(boards), L.ptr(y), B, st.bn_momentum, 0, L.ptr(st.loss), L.stream_ptr(st.device)))
        L.check(L.load_library().b2048_train_adam(L.ctx(st.device), L.ptr(st.params), L.ptr(st.grad), L.ptr(st.m), L.ptr(st.v),
                st.lr, 0.9, 0.999, 1e-8, 0.0, 100, L.stream_ptr(st.device)))
us_g = timed(lambda: g.replay())
print(f"b200 fused step, CUDA graph replay: {us_g:.0f} us/step", flush=True)
stm = pkg.trainer.TrainStep(pkg.network.ConvNet(64, 3).cuda().train(), lr=1e-3, max_batch=B, precision="mixed_fp32_wgrad")
print(f"b200 fused step, mixed precision (tcgen05 fwd/dgrad convs), graph: {timed(lambda: stm.step(boards, y)):.0f} us/step", flush=True)
stw = pkg.trainer.TrainStep(pkg.network.ConvNet(64, 3).cuda().train(), lr=1e-3, max_batch=B, precision="mixed")
print(f"b200 fused step, mixed + tensor-core weight gradient, graph: {timed(lambda: stw.step(boards, y)):.0f} us/step", flush=True)
def fwd(m, x):
    x = m.in_block(x)
    for b in m.blocks: x = m.relu(b(x) + x)
    return m.soft(m.policy(m.out_block(x).reshape(-1, m.out_size)))
x = torch.nn.functional.one_hot((boards.view(-1, 1) >> (4 * torch.arange(16, device="cuda"))) & 0xF, 16).permute(0, 2, 1).reshape(-1, 16, 4, 4).float()
torch.backends.cudnn.benchmark = True
for tf32 in (False, True):
    torch.backends.cudnn.allow_tf32 = tf32; torch.backends.cuda.matmul.allow_tf32 = tf32
    r = pkg.network.ConvNet(64, 3).cuda().train()
    opt = torch.optim.Adam(r.parameters(), lr=1e-3); lf = nn.KLDivLoss(reduction='batchmean')
    def step():
        opt.zero_grad(); lf(fwd(r, x), y).backward(); opt.step()
    print(f"torch eager step bs={B} tf32={tf32}: {timed(step):.0f} us/step", flush=True)
