"""GPU: population evaluation (evolve.py:14-67).  The multi-model launch must give, for every network,
exactly the games of the single-model kernel -- which test_network_gpu.py pins to the oracle and the goldens."""
import numpy as np
import pytest
import torch

from conftest import state_dict_from_golden

pytestmark = pytest.mark.gpu


def _net(pkg, ng, tag, precision):
    m = pkg.network.ConvNet(channels=64, blocks=3, precision=precision)
    m.load_state_dict(state_dict_from_golden(ng, tag))
    return m.cuda().eval()


def _perturbed(pkg, base, k, scale):
    g = torch.Generator(device="cuda").manual_seed(100 + k)
    twin = pkg.evolve.copynet(base).eval()
    with torch.no_grad():
        for w in twin.parameters():
            w.add_(scale * torch.randn(w.shape, device=w.device, generator=g))
    return twin


@pytest.mark.parametrize("precision", ["fp16", "fp32"])
def test_population_rows_equal_single_model_games(pkg, ng, precision):
    base = _net(pkg, ng, "shipped", precision)
    nets = [base, _net(pkg, ng, "random", precision)] + [_perturbed(pkg, base, k, 0.02) for k in range(3)]
    number = 24 if precision == "fp16" else 6
    blobs = torch.stack([m.blob(precision) for m in nets]).contiguous()
    for stride in (0, number):  # common random numbers / independent spawn streams
        pop = pkg.engine.playout_nn_population(blobs, precision, number, seed=31, gid_base=5, gid_model_stride=stride)
        for i, m in enumerate(nets):
            one = pkg.engine.playout_nn(m.blob(precision), precision, number, 31, 5 + i * stride)
            for key in ("final", "score", "moves"):
                assert torch.equal(pop[key][i], one[key]), (precision, stride, i, key)
        assert int(pop["total_moves"].item()) == int(pop["moves"].sum().item())


def test_population_larger_than_sm_count_and_first_death_groups(pkg, ng):
    # more networks than SMs -> several launches; group_min = eval_nn_min's first-death step per batch of 4 games
    base = _net(pkg, ng, "random", "fp16")
    nets = [_perturbed(pkg, base, k, 0.05) for k in range(12)]
    sms = pkg.engine.num_sms()
    reps = (sms + 20 + len(nets) - 1) // len(nets)
    blobs = torch.stack([m.blob("fp16") for m in nets] * reps).contiguous()
    M, number = blobs.shape[0], 8
    assert M > sms
    pop = pkg.engine.playout_nn_population(blobs, "fp16", number, seed=77, group_size=4)
    full = pkg.engine.playout_nn_population(blobs, "fp16", number, seed=77)
    want = full["moves"].reshape(M, 2, 4).amin(dim=2)
    assert torch.equal(pop["group_min"], want)
    for i in (0, len(nets) - 1, M - 1):  # replicas of one network play identical games
        assert torch.equal(full["moves"][i], full["moves"][i % len(nets)])
        one = pkg.engine.playout_nn(nets[i % len(nets)].blob("fp16"), "fp16", number, 77, 0)
        assert torch.equal(full["score"][i], one["score"])


def test_network_population_eval_select(pkg, ng, capsys):
    torch.manual_seed(3)
    np.random.seed(3)
    base = _net(pkg, ng, "shipped", "fp16")
    nets = [base] + [_perturbed(pkg, base, k, 0.3) for k in range(9)]
    pop = pkg.evolve.NetworkPopulation(nets, eval_num=16, keep_ratio=5, seed=11, precision="fp16")
    pop.eval()
    assert "1 of 10" in capsys.readouterr().out
    assert np.all(np.diff(pop.evals) <= 0)  # best first
    # ranking agrees with eval_nn run model by model on the same spawn streams
    for net, e in zip(pop.networks, pop.evals):
        mean_log, _, _ = pkg.eval_nn.eval_nn(net, number=16, seed=11, precision="fp16")
        assert mean_log == pytest.approx(e, abs=1e-12)
    assert pop.networks[0] is base or pop.evals[0] >= pkg.eval_nn.eval_nn(base, number=16, seed=11, precision="fp16")[0]
    nxt = pop.select()
    assert len(nxt) == 10 and nxt[0] is pop.networks[0] and nxt[1] is pop.networks[1]
    for child in nxt[2:]:
        assert sum(p.numel() for p in child.parameters()) == 231820
        assert child is not pop.networks[0]
    # a mutation changes about mutation_rate of the elements; a recombination takes whole tensors from either parent
    kid = pop.mutate(base)
    changed = sum(int((a != b).sum()) for a, b in zip(kid.parameters(), base.parameters()))
    assert 0.005 * 231820 < changed < 0.02 * 231820
    other = nets[1]
    mix = pop.recombine(base, other)
    for w, a, b in zip(mix.parameters(), base.parameters(), other.parameters()):
        assert torch.equal(w, a) or torch.equal(w, b)
