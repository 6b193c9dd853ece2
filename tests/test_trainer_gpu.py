"""GPU: the hand-written fp32 training step (csrc/train.cuh) against PyTorch autograd on the same module tree
(trainer.py:55-62: train-mode forward, KLDivLoss(batchmean), backward, Adam)."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def dev_boards(x):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x, np.uint64).view(np.int64))).cuda()


def torch_forward(m, x, margins=None):
    """network.py:78-87 with plain torch modules (our ConvNet.forward dispatches to the CUDA inference kernels).
    margins: list collecting min |pre-activation| of every ReLU (distance of the batch from a ReLU branch flip)."""
    def act(t):
        if margins is not None:
            margins.append(t.detach().abs().min().item())
        return torch.relu(t)
    x = act(m.in_block[1](m.in_block[0](x)))
    for blk in m.blocks:
        h = act(blk[1](blk[0](x)))
        x = act(blk[4](blk[3](h)) + x)
    h = act(m.out_block[1](m.out_block[0](x)))
    return m.soft(m.policy(h.reshape(-1, m.out_size)))


def onehot(boards):
    expo = (boards.view(-1, 1) >> (4 * torch.arange(16, device=boards.device))) & 0xF  # (N,16) position-major
    return torch.nn.functional.one_hot(expo, 16).permute(0, 2, 1).reshape(-1, 16, 4, 4)


def make_pair(pkg, seed, dtype):
    torch.manual_seed(seed)
    ours = pkg.network.ConvNet(64, 3).cuda()
    with torch.no_grad():  # non-trivial BatchNorm affine parameters and running statistics
        for mod in ours.modules():
            if isinstance(mod, nn.BatchNorm2d):
                mod.weight.uniform_(0.5, 1.5), mod.bias.uniform_(-0.3, 0.3)
                mod.running_mean.uniform_(-0.2, 0.2), mod.running_var.uniform_(0.5, 2.0)
    ref = pkg.network.ConvNet(64, 3).cuda()
    ref.load_state_dict(ours.state_dict())
    return ours.train(), ref.to(dtype).train()


def batch(ng, B, seed):
    g = torch.Generator().manual_seed(seed)
    idx = torch.randint(0, len(ng["lg_boards"]), (B,), generator=g).numpy()
    boards = dev_boards(ng["lg_boards"][idx])
    y = torch.softmax(3.0 * torch.rand(B, 4, generator=g), dim=1).cuda()
    y[::7, 1] = 0.0  # exact zeros in the targets: xlogy(0, 0) = 0 in KLDivLoss
    y = y / y.sum(1, keepdim=True)
    return boards, y.contiguous()


@pytest.fixture()
def true_fp32():
    prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev


def _grad_errors(stepper, ref):
    off, out = 0, {}
    for name, p in ref.named_parameters():
        n = p.numel()
        mine = stepper.grad[off:off + n].view(p.shape).double()
        out[name] = ((mine - p.grad).abs().max() / p.grad.abs().max()).item()
        off += n
    assert off == 231820
    return out


@pytest.mark.parametrize("B", [64, 50])
def test_gradients_match_fp64_autograd_exactly_away_from_relu_flips(pkg, ng, B, true_fp32):
    # An activation within rounding of 0 takes different ReLU branches in fp32 and fp64 and changes every gradient
    # below it by O(1/B) (PyTorch's own fp32 autograd shows the same ~1e-3 jumps).  Pick a batch whose smallest
    # |pre-activation| is well clear of rounding, then every gradient tensor must match fp64 autograd to 2e-5.
    for seed in range(2, 80):
        ours, ref = make_pair(pkg, 1, torch.float64)
        boards, y = batch(ng, B, seed)
        margins = []
        ref_logp = torch_forward(ref, onehot(boards).double(), margins)
        if min(margins) > 4e-6:
            break
    else:
        pytest.skip("no batch without near-threshold activations found")
    ref_loss = nn.KLDivLoss(reduction='batchmean')(ref_logp, y.double())
    ref_loss.backward()
    stepper = pkg.trainer.TrainStep(ours, max_batch=64)
    loss = float(stepper.forward_backward(boards, y).item())
    assert loss == pytest.approx(float(ref_loss.item()), rel=1e-5)
    assert (stepper.log_probs(B).double() - ref_logp.detach()).abs().max() < 1e-5
    for name, err in _grad_errors(stepper, ref).items():
        assert err <= 2e-5, (name, err, seed)
    for (name, b_ref), (_, b_ours) in zip(ref.named_buffers(), ours.named_buffers()):
        if name.endswith("num_batches_tracked"):
            assert int(b_ours) == int(b_ref) == 1
        else:  # running statistics: momentum 0.1, unbiased batch variance
            assert torch.allclose(b_ours.double(), b_ref, rtol=1e-5, atol=1e-6), name


@pytest.mark.parametrize("B", [250, 2048])
def test_full_batch_loss_and_gradients(pkg, ng, B, true_fp32):
    # trainer.py's batch size: ReLU branch flips are unavoidable among 15 M activations, so the gradient bound is
    # the one PyTorch's fp32 autograd meets against fp64 (a few 1e-3 of the tensor's scale in the lower layers);
    # loss, log-probs and the layers above the lowest flip stay tight.
    ours, ref = make_pair(pkg, 1, torch.float64)
    _, ref32 = make_pair(pkg, 1, torch.float32)
    boards, y = batch(ng, B, 2)
    stepper = pkg.trainer.TrainStep(ours, max_batch=2048)
    loss = float(stepper.forward_backward(boards, y).item())
    ref_logp = torch_forward(ref, onehot(boards).double())
    ref_loss = nn.KLDivLoss(reduction='batchmean')(ref_logp, y.double())
    ref_loss.backward()
    nn.KLDivLoss(reduction='batchmean')(torch_forward(ref32, onehot(boards).float()), y).backward()
    assert loss == pytest.approx(float(ref_loss.item()), rel=2e-5)
    assert (stepper.log_probs(B).double() - ref_logp.detach()).abs().max() < 2e-5
    errs = _grad_errors(stepper, ref)
    t32 = {n: ((q.grad.double() - p.grad).abs().max() / p.grad.abs().max()).item()
           for (n, p), (_, q) in zip(ref.named_parameters(), ref32.named_parameters())}
    assert max(errs.values()) <= max(3 * max(t32.values()), 2e-2), (errs, t32)
    for name in ("policy.weight", "policy.bias", "out_block.0.weight", "out_block.1.weight", "out_block.1.bias"):
        assert errs[name] <= 2e-5, (name, errs[name])


def test_adam_update_matches_torch_optim(pkg):
    torch.manual_seed(0)
    stepper = pkg.trainer.TrainStep(pkg.network.ConvNet(64, 3).cuda(), lr=0.1, weight_decay=1e-3, max_batch=8)
    ref = nn.Parameter(stepper.params.clone())
    opt = torch.optim.Adam([ref], lr=0.1, weight_decay=1e-3)
    for t in range(4):
        g = torch.randn(231820, device="cuda") * 10.0 ** torch.randint(-9, 1, (231820,), device="cuda").float()
        stepper.adam(g)
        ref.grad = g.clone()
        opt.step()
        assert (stepper.params - ref.detach()).abs().max() <= 1e-5, t  # 1e-4 of one lr-sized update
    m_ref, v_ref = opt.state[ref]["exp_avg"], opt.state[ref]["exp_avg_sq"]
    assert torch.allclose(stepper.m, m_ref, rtol=1e-5, atol=1e-6 * m_ref.abs().max().item())  # sums can cancel
    assert torch.allclose(stepper.v, v_ref, rtol=1e-5, atol=1e-30)


def test_adam_trajectory_matches_torch_optim(pkg, ng, true_fp32):
    # three runs of the same 6 steps: ours, torch fp32, torch fp64.  Adam's normalised update turns ReLU branch flips
    # and the rounding of small gradients into O(lr) parameter differences, so every fp32 run drifts from fp64;
    # ours must stay in the same order of drift as PyTorch's own fp32 run, and the losses must agree.
    ours, ref64 = make_pair(pkg, 5, torch.float64)
    _, ref32 = make_pair(pkg, 5, torch.float32)
    stepper = pkg.trainer.TrainStep(ours, lr=1e-3, weight_decay=1e-4, max_batch=512)
    opts = [torch.optim.Adam(r.parameters(), lr=1e-3, weight_decay=1e-4) for r in (ref64, ref32)]
    loss_fn = nn.KLDivLoss(reduction='batchmean')
    for it in range(6):
        boards, y = batch(ng, 512, 10 + it)
        mine = float(stepper.step(boards, y).item())
        for r, opt, dt in ((ref64, opts[0], torch.float64), (ref32, opts[1], torch.float32)):
            opt.zero_grad()
            loss = loss_fn(torch_forward(r, onehot(boards).to(dt)), y.to(dt))
            loss.backward()
            opt.step()
            assert mine == pytest.approx(float(loss.item()), rel=2e-3), (it, dt)
    flat = lambda r: torch.cat([p.detach().reshape(-1).double() for p in r.parameters()])
    d_ours = (stepper.params.double() - flat(ref64)).abs()
    d_t32 = (flat(ref32) - flat(ref64)).abs()
    assert d_ours.mean() <= 10 * d_t32.mean() + 1e-6  # same order of drift as PyTorch's fp32 run (ReLU flips differ)
    assert d_ours.max() < 6 * 2 * 1e-3
    # the module's tensors are views of the flat vectors the kernels update
    assert torch.equal(torch.cat([p.detach().reshape(-1) for p in ours.parameters()]), stepper.params)
    assert int(ours.in_block[1].num_batches_tracked) == 6


def test_step_is_deterministic_and_updates_inference_weights(pkg, ng):
    outs = []
    for _ in range(2):
        ours, _ = make_pair(pkg, 9, torch.float32)
        stepper = pkg.trainer.TrainStep(ours, lr=0.01, max_batch=1024)
        for it in range(3):
            boards, y = batch(ng, 1000, 40 + it)
            stepper.step(boards, y)
        outs.append((stepper.params.clone(), stepper.bn.clone(), float(stepper.loss.item())))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and outs[0][2] == outs[1][2]
    # eval-mode inference through the playout kernels sees the trained weights (weight image cache invalidated)
    ours.eval()
    boards, _ = batch(ng, 512, 99)
    lp = ours.forward_boards(boards, precision="fp32")
    with torch.no_grad():
        ref = torch_forward(ours, onehot(boards).float())
    assert (lp - ref).abs().max() < 1e-4


def test_trainer_main_loop(pkg, ng, tmp_path, capsys):
    # records in the reference's format (selfplay.py:55-58), a short training run, checkpoint + log written
    rng = np.random.default_rng(0)
    os.makedirs(tmp_path / "selfplay")
    os.makedirs(tmp_path / "models")
    os.makedirs(tmp_path / "logs")
    for i in range(4):
        idx = rng.integers(0, len(ng["lg_boards"]), 300)
        results = rng.uniform(50, 400, (300, 4)).astype(np.float32)
        np.savez(tmp_path / "selfplay" / (str(i).zfill(5) + ".npz"), boards=ng["lg_boards"][idx], results=results, score=1000)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        args = pkg.trainer.parser().parse_args(["--path", "selfplay/", "--t_tuple", "0", "4", "--epochs", "2", "--batch_size",
                                                "256", "--lr", "0.01", "--patience", "5", "--name", "t"])
        t_loss, min_move = pkg.trainer.main(args)
    finally:
        os.chdir(cwd)
    assert len(t_loss) == 2 and len(min_move) == 2 and all(np.isfinite(t_loss)) and t_loss[1] < t_loss[0]
    saved = [f for f in os.listdir(tmp_path / "models") if f.endswith("_best.pt")]
    assert len(saved) == 1
    sd = torch.load(tmp_path / "models" / saved[0])
    assert len(sd) == 50  # the reference's checkpoint keys
    assert os.listdir(tmp_path / "logs")
    assert "Train mLoss" in capsys.readouterr().out


@pytest.mark.parametrize("B,mode", [(250, "mixed"), (2048, "mixed"), (250, "mixed_fp32_wgrad"), (32, "mixed"), (33, "mixed")])
def test_mixed_precision_step_tracks_fp64(pkg, ng, B, mode):
    # precision="mixed": forward / data-gradient / weight-gradient convolutions on the tensor cores (fp16 / bf16 operands,
    # fp32 accumulate; the weight gradient uses MN-major operands and sums over whole 32-board tiles -> ragged B checked).
    # Bar = autocast-like agreement with fp64 autograd: loss 1e-3, log-probs 5e-3, every gradient tensor within 10 % in
    # L2 and cosine >= 0.995 (measured: loss 1.4e-5, log-probs 5e-4, 1.5-4 %, >= 0.9991; most of it ReLU branch flips
    # of activations that fp16 moves across 0)
    ours, ref = make_pair(pkg, 1, torch.float64)
    boards, y = batch(ng, B, 2)
    stepper = pkg.trainer.TrainStep(ours, max_batch=2048, precision=mode)
    stepper.workspace.fill_(0xFF)  # stale / NaN bit patterns everywhere: rows past the batch end must not leak in
    loss = float(stepper.forward_backward(boards, y).item())
    ref_logp = torch_forward(ref, onehot(boards).double())
    ref_loss = nn.KLDivLoss(reduction='batchmean')(ref_logp, y.double())
    ref_loss.backward()
    assert loss == pytest.approx(float(ref_loss.item()), rel=1e-3)
    assert (stepper.log_probs(B).double() - ref_logp.detach()).abs().max() < 5e-3
    off = 0
    for name, p in ref.named_parameters():
        n = p.numel()
        mine = stepper.grad[off:off + n].view(p.shape).double()
        off += n
        rel = float((mine - p.grad).norm() / p.grad.norm())
        cos = float((mine * p.grad).sum() / (mine.norm() * p.grad.norm()))
        assert rel < (0.1 if B >= 250 else 0.2) and cos > (0.995 if B >= 250 else 0.98), (name, rel, cos)


def test_mixed_precision_training_converges_like_fp32(pkg, ng):
    # the same 30 Adam steps on a fixed synthetic task in both modes: same loss curve within a few percent, deterministic
    losses = {}
    for prec in ("fp32", "mixed", "mixed"):
        ours, _ = make_pair(pkg, 3, torch.float32)
        stepper = pkg.trainer.TrainStep(ours, lr=2e-3, max_batch=512, precision=prec)
        out = []
        for it in range(30):
            boards, y = batch(ng, 512, 100 + (it % 5))
            out.append(float(stepper.step(boards, y).item()))
        losses.setdefault(prec, []).append(out)
    fp32, mixed, again = losses["fp32"][0], losses["mixed"][0], losses["mixed"][1]
    assert mixed == again  # bit-deterministic
    assert fp32[-1] < 0.7 * fp32[0] and mixed[-1] < 0.7 * mixed[0]
    assert abs(mixed[-1] - fp32[-1]) < 0.05 * fp32[-1] + 1e-3
    assert max(abs(a - b) for a, b in zip(fp32[:5], mixed[:5])) < 2e-3
