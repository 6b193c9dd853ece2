"""Drop-in for the reference's `trainer` module (fqjin/2048NN trainer.py:14-97): supervised training of c64b3 on
self-play records with KLDivLoss(batchmean) + Adam, evaluated every epoch by `eval_nn_min` (first-death step of
policy games), best-only checkpoints, patience-based early stopping.

The optimisation step is NOT autograd: `TrainStep` drives the hand-written fp32 CUDA training kernels
(b2048_train_forward_backward + b2048_train_adam, csrc/train.cuh) on packed uint64 boards; the module's parameters
and BatchNorm buffers alias flat device vectors the kernels update in place, so `state_dict()` / checkpoints
keep the reference's 50 keys.  The in-loop evaluation is the tensor-core playout kernel (K4).
"""
import argparse
import random
from time import time

import numpy as np
import torch

from . import _lib as L
from .eval_nn import eval_nn_min
from .game_dataset import OneHotConvGameDataset
from .network import ConvNet

N_PARAMS = 231820
_BN_NAMES = ["in_block.1"] + [f"blocks.{k}.{j}" for k in range(3) for j in (1, 4)] + ["out_block.1"]


class TrainStep:
    """Binds a c64b3 `ConvNet` to the fused training step.

    step(boards, targets) == trainer.py:56-62 (zero_grad, forward in train mode, KLDivLoss(batchmean), backward,
    Adam step) for one batch: boards int64 (uint64 bit patterns) [B], targets float32 [B,4].  Returns the loss as a
    0-d device tensor (no host sync).  precision="mixed" runs the forward / data-gradient / weight-gradient convolutions
    on the tensor cores (fp16 / bf16 operands, fp32 accumulation; BatchNorm, head, Adam and the master weights fp32) --
    about twice as fast, autocast-like numerics."""

    def __init__(self, model, lr=0.1, weight_decay=0.0, betas=(0.9, 0.999), eps=1e-8, max_batch=2048, bn_momentum=0.1,
                 use_graph=True, precision="fp32"):
        L.require_cuda()
        if getattr(model, "_shape", (64, 3, 4)) != (64, 3, 4):
            raise NotImplementedError("the B200 training kernels implement c64b3 = ConvNet(channels=64, blocks=3)")
        params = list(model.parameters())
        if sum(p.numel() for p in params) != N_PARAMS:
            raise ValueError("model is not a c64b3 ConvNet (231,820 parameters)")
        self.model, self.lr, self.wd, self.betas, self.eps = model, float(lr), float(weight_decay), betas, float(eps)
        self.max_batch, self.bn_momentum = int(max_batch), float(bn_momentum)
        modes = {"fp32": 0, "mixed_fp32_wgrad": 1, "mixed": 2}
        if precision not in modes:
            raise ValueError('precision: "fp32" (reference arithmetic), "mixed" (all 3x3 convolutions -- forward, data gradient, '
                             'weight gradient -- on the tensor cores) or "mixed_fp32_wgrad" (weight gradient kept in fp32)')
        self.mixed = modes[precision]
        dev = params[0].device if params[0].device.type == "cuda" else torch.device("cuda")
        self.device = dev
        # flat parameter / buffer vectors; the module's tensors become views of them
        self.params = torch.empty(N_PARAMS, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in params:
                n = p.numel()
                self.params[off:off + n].copy_(p.detach().reshape(-1))
                p.data = self.params[off:off + n].view(p.shape)
                off += n
            self.bn = torch.zeros(2 * 8 * 64, dtype=torch.float32, device=dev)
            mods = dict(model.named_modules())
            self._bn_mods = [mods[n] for n in _BN_NAMES]
            for i, bn in enumerate(self._bn_mods):
                c = bn.running_mean.numel()
                self.bn[i * 64:i * 64 + c].copy_(bn.running_mean)
                self.bn[512 + i * 64:512 + i * 64 + c].copy_(bn.running_var)
                bn.running_mean = self.bn[i * 64:i * 64 + c]
                bn.running_var = self.bn[512 + i * 64:512 + i * 64 + c]
            self._nbt = torch.stack([bn.num_batches_tracked.to(dev) for bn in self._bn_mods]).contiguous()
            for i, bn in enumerate(self._bn_mods):
                bn.num_batches_tracked = self._nbt[i]
        self.m = torch.zeros(N_PARAMS, dtype=torch.float32, device=dev)
        self.v = torch.zeros(N_PARAMS, dtype=torch.float32, device=dev)
        lib = L.load_library()
        nbytes = lib.b2048_train_workspace_bytes(self.max_batch)
        self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        g0 = lib.b2048_train_grad_offset(self.max_batch)
        self.grad = self.workspace[g0:g0 + 4 * N_PARAMS].view(torch.float32)
        l0 = lib.b2048_train_logp_offset(self.max_batch)
        self._logp = self.workspace[l0:l0 + 16 * self.max_batch].view(torch.float32).view(-1, 4)
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.steps = 0
        # the ~65 launches of forward+backward replay as one CUDA graph per batch size (static input buffers)
        self.use_graph = bool(use_graph)
        self._graphs = {}
        self._in_boards = torch.zeros(self.max_batch, dtype=torch.int64, device=dev)
        self._in_targets = torch.zeros((self.max_batch, 4), dtype=torch.float32, device=dev)

    def forward_backward(self, boards, targets):
        """Loss and gradient of one batch (BatchNorm running statistics are updated as in train mode)."""
        B = boards.numel()
        boards, targets = boards.contiguous(), targets.contiguous()
        if targets.shape != (B, 4) or targets.dtype != torch.float32 or boards.dtype != torch.int64:
            raise ValueError("boards: int64 [B], targets: float32 [B,4]")
        if B > self.max_batch:
            raise ValueError("batch larger than max_batch")
        first = next(self.model.parameters())
        if first.data_ptr() != self.params.data_ptr():
            raise RuntimeError("the model's parameters no longer alias the trainer's flat buffer (model.to()/.cpu() after "
                               "TrainStep was built?): build a new TrainStep for the moved model")
        if not self.use_graph:
            self._enqueue(boards, targets, B)
        else:
            self._in_boards[:B].copy_(boards)
            self._in_targets[:B].copy_(targets)
            g = self._graphs.get(B)
            if g is None:
                self._enqueue(self._in_boards, self._in_targets, B)  # eager once: also the capture warm-up
                torch.cuda.synchronize(self.device)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._enqueue(self._in_boards, self._in_targets, B)
                self._graphs[B] = g  # (capturing records the launches, it does not run them)
            else:
                g.replay()
        self._nbt += 1
        self._invalidate_weight_images()
        return self.loss[0]

    def _invalidate_weight_images(self):
        # the kernels write parameters / running statistics behind autograd's version counters: drop the cached
        # inference weight images (network.ConvNet.blob, network.blob_for) so the next playout re-folds them
        self.model.__dict__["_blob_key"] = None
        self.model.__dict__.pop("_b2048_blobs", None)

    def _enqueue(self, boards, targets, B):
        L.check(L.load_library().b2048_train_forward_backward(
            L.ctx(self.device), L.ptr(self.params), L.ptr(self.bn), L.ptr(self.workspace), self.max_batch, L.ptr(boards),
            L.ptr(targets), B, self.bn_momentum, self.mixed, L.ptr(self.loss), L.stream_ptr(self.device)))

    def log_probs(self, B):
        """log-probabilities the last forward_backward computed for its batch (train-mode forward)."""
        return self._logp[:B]

    def adam(self, grad=None):
        self.steps += 1
        g = self.grad if grad is None else grad
        L.check(L.load_library().b2048_train_adam(L.ctx(self.device), L.ptr(self.params), L.ptr(g), L.ptr(self.m), L.ptr(self.v),
                                                  self.lr, self.betas[0], self.betas[1], self.eps, self.wd, self.steps,
                                                  L.stream_ptr(self.device)))
        self._invalidate_weight_images()

    def step(self, boards, targets, group=None):
        """One optimisation step.  With an initialised process group it is data parallel: every rank runs its own
        batch and the flat gradient is averaged over the ranks between backward and Adam (one allreduce), so the
        parameters stay identical on all ranks.  BatchNorm batch statistics stay per rank, as in
        DistributedDataParallel without SyncBatchNorm; the running statistics are averaged by sync_buffers()."""
        loss = self.forward_backward(boards, targets)
        td = torch.distributed
        if td.is_available() and td.is_initialized() and td.get_world_size(group) > 1:
            td.all_reduce(self.grad, group=group)
            self.grad /= td.get_world_size(group)
        self.adam()
        return loss

    def sync_buffers(self, group=None):
        """Average the BatchNorm running statistics over the ranks.  The running-statistics update is linear in
        (running value, batch statistic), so averaging once -- before an evaluation or a checkpoint -- gives exactly
        the buffers that averaging after every step would: no per-step collective is needed for them."""
        td = torch.distributed
        if td.is_available() and td.is_initialized() and td.get_world_size(group) > 1:
            td.all_reduce(self.bn, group=group)
            self.bn /= td.get_world_size(group)
            self._invalidate_weight_images()


def _world(group=None):
    td = torch.distributed
    if td.is_available() and td.is_initialized():
        return td.get_rank(group), td.get_world_size(group)
    return 0, 1


def train_epoch(stepper, boards, targets, batch_size, generator=None, group=None):
    """One shuffled pass (DataLoader(shuffle=True), trainer.py:29,55): returns (mean batch loss, batches).
    Data parallel: every rank draws the SAME permutation (pass generators seeded alike, or rely on torch.manual_seed
    having been called with the same seed on every rank) and takes its own `batch_size` slice of each global batch of
    world_size * batch_size records, so an epoch still visits every record once."""
    rank, world = _world(group)
    n = boards.numel()
    perm = torch.randperm(n, device=boards.device, generator=generator)
    if world > 1:
        torch.distributed.broadcast(perm, 0, group=group)  # one permutation for all ranks, whatever their RNG state
    total = torch.zeros((), dtype=torch.float32, device=boards.device)
    batches = 0
    for i in range(0, n, batch_size * world):
        idx = perm[i + rank * batch_size:i + (rank + 1) * batch_size]
        short = torch.tensor([idx.numel() < 2], device=boards.device)
        if world > 1:
            torch.distributed.all_reduce(short, op=torch.distributed.ReduceOp.MAX, group=group)
        if bool(short.item()):
            break  # a ragged last global batch that leaves some rank without records is dropped on every rank
        total += stepper.step(boards[idx], targets[idx], group=group)
        batches += 1
    return float(total.item()) / max(batches, 1), batches


class Patience:
    """Early stopping of trainer.py:46-52,66-67,81-91: stop after `patience` epochs without a new best evaluation
    (patience 0: never), and give up at once if the third epoch's mean loss is still above 0.210."""

    def __init__(self, patience, epochs):
        self.limit = epochs if patience == 0 else patience
        self.best, self.since_best = 0.0, 0

    def after_train(self, epoch, mean_loss):
        self.since_best += 1
        if epoch == 2 and mean_loss > 210 / 1000:
            self.limit = 0

    def after_eval(self, value):
        """-> True when `value` is a new best (ties count, trainer.py:81)."""
        if value >= self.best:
            self.best, self.since_best = value, 0
            return True
        return False

    @property
    def exhausted(self):
        return self.since_best >= self.limit

    @property
    def remaining(self):
        return self.limit - self.since_best


def run_name(args, pretrained):
    """The reference's log / checkpoint name grammar (trainer.py:17-21,38)."""
    prefix = (args.name + '_') if args.name else ''
    name = (f'{prefix}{args.t_tuple[0]}_{args.t_tuple[1]}_soft{args.soft}c{args.channels}b{args.blocks}_p{args.patience}_'
            f'bs{args.batch_size}lr{args.lr}d{args.decay}_s{args.seed}')
    return 'pre_' + name if pretrained else name


def main(args, group=None):
    """trainer.py:14-97: train on the self-play records under args.path, evaluate with eval_nn_min after every epoch,
    keep the best checkpoint, stop on patience; returns (losses, evaluations).  Under torchrun the run is data parallel
    (see TrainStep.step / train_epoch); only rank 0 prints, evaluates and writes models/ and logs/."""
    if (args.channels, args.blocks) != (64, 3):
        raise SystemExit('the B200 kernels implement c64b3 only: use --channels 64 --blocks 3 '
                         f'(got c{args.channels}b{args.blocks})')
    rank, world = _world(group)
    say = print if rank == 0 else (lambda *a, **k: None)
    logname = run_name(args, bool(args.pretrained))
    say(logname)
    for seeder in (random.seed, np.random.seed, torch.manual_seed):
        seeder(args.seed)
    L.require_cuda()
    device = torch.device('cuda')
    say('Using {}'.format(device))

    records = OneHotConvGameDataset(args.path, args.t_tuple[0], args.t_tuple[1], device, soft=args.soft)
    net = ConvNet(channels=64, blocks=3, precision="fp16")
    if args.pretrained:
        net.load_state_dict(torch.load('models/{}.pt'.format(args.pretrained), map_location=device))
        say('Loaded ' + args.pretrained)
    net.to(device)
    stepper = TrainStep(net, lr=args.lr, weight_decay=args.decay, max_batch=args.batch_size,
                        precision=getattr(args, "precision", "fp32"))
    if world > 1:  # same start on every rank (DDP broadcasts parameters and buffers once at construction)
        torch.distributed.broadcast(stepper.params, 0, group=group)
        torch.distributed.broadcast(stepper.bn, 0, group=group)
    patience = Patience(args.patience, args.epochs)
    losses, evals = [], []
    for epoch in range(args.epochs):
        say('-' * 10)
        say('Epoch: {}'.format(epoch))
        net.train()
        mean_loss, _ = train_epoch(stepper, records.packed, records.results, args.batch_size, group=group)
        stepper.sync_buffers(group)  # identical BatchNorm buffers on every rank before evaluating / saving
        patience.after_train(epoch, mean_loss)
        say('Train mLoss: {:.3f}'.format(1e3 * mean_loss))
        losses.append(mean_loss)

        net.eval()
        started = time()
        # every rank plays the same evaluation games (same seed, identical weights): the decision below is identical
        # everywhere without an exchange
        first_death = eval_nn_min(net, number=10, repeats=40, device=device)
        took = ', took {:.0f} seconds'.format(time() - started)
        evals.append(first_death)
        if patience.after_eval(first_death):
            say(str(first_death) + ' ** Best' + took)
            if rank == 0:
                torch.save(net.state_dict(), 'models/' + logname + '_best.pt')
        else:
            say(str(first_death) + took)
        if patience.exhausted:
            say('Ran out of patience')
            say(f'Best score: {patience.best}')
            break
        say(f'{patience.remaining} epochs remaining')
    if rank == 0:
        np.savez('logs/' + logname, t_loss=losses, min_move=evals, params=args)
    return losses, evals


def parser():
    p = argparse.ArgumentParser()
    p.add_argument('--path', type=str, default='selfplay/', help='path to training data with prefix')
    p.add_argument('--t_tuple', type=int, nargs=2, default=(20, 200), help='tuple for training data range')
    p.add_argument('--channels', type=int, default=64)
    p.add_argument('--blocks', type=int, default=3)
    p.add_argument('--epochs', type=int, default=1000)
    p.add_argument('--patience', type=int, default=10,
                   help='Early stopping based on log score eval. If zero, no early stopping.')
    p.add_argument('--batch_size', type=int, default=2048)
    p.add_argument('--lr', type=float, default=0.1)
    p.add_argument('--decay', type=float, default=0.0)
    p.add_argument('--soft', type=float, default=3.0)
    p.add_argument('--name', type=str, default='', help='Additional prepend output name')
    p.add_argument('--seed', type=int, default=0)
    p.add_argument('--pretrained', type=str, default='', help='Path to network to continue training')
    p.add_argument('--precision', type=str, default='fp32', choices=['fp32', 'mixed'],
                   help='fp32 = the reference arithmetic; mixed = convolutions on the tensor cores (not in the reference)')
    return p


if __name__ == '__main__':
    main(parser().parse_args())
