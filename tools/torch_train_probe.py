"""How long does the reference's own training step (trainer.py:55-62: PyTorch eager, cuDNN) take on this GPU?"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
pkg = importlib.import_module("2048nn_b200")
torch.backends.cudnn.benchmark = True
for tf32 in (True, False):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    m = pkg.network.ConvNet(64, 3).cuda().train()
    m.forward = lambda x, m=m: m.soft(m.policy(m.out_block(_body(m, x)).view(-1, m.out_size)))
    def _body(m, x):
        x = m.in_block(x)
        for b in m.blocks:
            x = m.relu(b(x) + x)
        return x
    opt = torch.optim.Adam(m.parameters(), lr=0.1)
    loss_fn = nn.KLDivLoss(reduction='batchmean')
    B = 2048
    x = torch.zeros(B, 16, 4, 4, device="cuda"); idx = torch.randint(0, 16, (B, 1, 4, 4), device="cuda"); x.scatter_(1, idx, 1.0)
    y = torch.softmax(torch.randn(B, 4, device="cuda"), 1)
    def step():
        opt.zero_grad(); loss = loss_fn(m(x), y); loss.backward(); opt.step(); return loss
    for _ in range(10): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): step()
    e1.record(); torch.cuda.synchronize()
    print(f"torch eager train step bs={B} tf32={tf32}: {e0.elapsed_time(e1)/50*1e3:.0f} us/step", flush=True)
