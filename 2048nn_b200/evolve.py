"""Drop-in for the reference's `evolve` module (fqjin/2048NN evolve.py): a population of c64b3 networks ranked
by policy-game evaluation, then truncation selection with mutation / recombination.

The reference evaluates one network after another (`eval_nn(model, number=eval_num)`, evolve.py:30-36): 100
concurrent games use about 1 % of a B200.  Here `eval()` packs every network's folded weight image into one
device buffer and plays all networks' games side by side in one persistent-kernel launch
(b2048_playout_nn_population); each CTA is bound to one network.  `select`, `mutate`, `recombine` act on
parameters, not on the playout path, and keep the reference's semantics (evolve.py:43-67).
"""
import random

import numpy as np
import torch

from . import engine
from .board import _raise_if_flag, _session_seed
from .network import ConvNet, blob_for


def copynet(net):
    """evolve.py:8-11: an independent c64b3 with the same parameters and buffers."""
    twin = ConvNet(channels=64, blocks=3, precision=getattr(net, "precision", "fp32"))
    twin.load_state_dict(net.state_dict())
    return twin.to(next(net.parameters()).device)


def eval_population(models, number=100, seed=None, precision=None, common_random_numbers=True):
    """eval_nn (eval_nn.py:4-58) for every model at once -> list of (mean log10(score+1), max score, mean moves).
    With common_random_numbers every model plays the same `number` spawn streams (game ids 0..number-1, exactly
    what eval_nn(model, number, seed=seed) plays), which removes the spawn noise from comparisons between models."""
    packed = [blob_for(m, precision) for m in models]
    precs = {p for _, p in packed}
    if len(precs) != 1:
        raise ValueError("all networks of a population must use the same precision")
    blobs = torch.stack([b for b, _ in packed]).contiguous()
    r = engine.playout_nn_population(blobs, precs.pop(), int(number), _session_seed(seed), 0,
                                     0 if common_random_numbers else int(number), want_final=False)
    _raise_if_flag()
    scores = r["score"].cpu().numpy().view(np.uint32).astype(np.int64)
    moves = r["moves"].cpu().numpy().astype(np.int64)
    return [(np.mean(np.log10(s + 1)), np.amax(s), np.mean(m)) for s, m in zip(scores, moves)]


class NetworkPopulation:
    """evolve.py:14-67."""

    def __init__(self, networks, eval_num=100, keep_ratio=5, recombine_prob=0.2, mutation_rate=0.01, seed=None,
                 precision=None):
        if isinstance(networks, list):
            arr = np.empty(len(networks), dtype=object)
            arr[:] = networks
            networks = arr
        self.networks = networks
        self.popsize = len(networks)
        self.keep_ratio = keep_ratio
        self.recombine_prob = recombine_prob
        self.mutation_rate = mutation_rate
        assert self.popsize >= keep_ratio
        self.eval_num = eval_num
        self.device = next(networks[0].parameters()).device
        self.evals = np.zeros(self.popsize)
        self.seed, self.precision = seed, precision

    def eval(self):
        """Rank the population by mean log10 score of eval_num policy games each (evolve.py:30-41)."""
        stats = eval_population(list(self.networks), self.eval_num, self.seed, self.precision)
        for i, (mean_log_score, max_score, mean_moves) in enumerate(stats):
            print(f'{i+1} of {self.popsize}')
            print(mean_log_score, max_score, mean_moves)
            self.evals[i] = mean_log_score
        best_first = np.argsort(-1 * self.evals)
        self.networks = self.networks[best_first]
        self.evals = self.evals[best_first]

    def select(self):
        """Keep the top popsize // keep_ratio networks; refill with recombinations (probability recombine_prob,
        two distinct parents) or mutations of random survivors (evolve.py:43-53)."""
        elite = self.networks[:(self.popsize // self.keep_ratio)]
        children = []
        for u in np.random.rand(self.popsize - len(elite)):
            if u < self.recombine_prob:
                children.append(self.recombine(*np.random.choice(elite, 2, replace=False)))
            else:
                children.append(self.mutate(np.random.choice(elite, 1).item()))
        out = np.empty(self.popsize, dtype=object)
        out[:len(elite)] = elite
        out[len(elite):] = children
        return out

    def mutate(self, network):
        """Each parameter element gets N(0,1) noise with probability mutation_rate (evolve.py:55-59)."""
        child = copynet(network)
        with torch.no_grad():
            for w in child.parameters():
                hit = torch.rand_like(w) < self.mutation_rate
                w.add_(hit * torch.randn_like(w))
        return child

    def recombine(self, network1, network2):
        """Per parameter tensor, a fair coin picks the parent it is taken from (evolve.py:61-67)."""
        child, other = copynet(network1), copynet(network2)
        with torch.no_grad():
            for a, b in zip(child.parameters(), other.parameters()):
                if random.getrandbits(1):
                    a.copy_(b)
        return child
