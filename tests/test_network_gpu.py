"""GPU: c64b3 forward and NN-policy playouts against the fp32 PyTorch oracle and the goldens
produced by the reference's own network.ConvNet / mcts_nn / eval_nn loops."""
import numpy as np
import pytest
import torch

from conftest import state_dict_from_golden

pytestmark = pytest.mark.gpu


def dev_boards(x):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x, np.uint64).view(np.int64))).cuda()


@pytest.fixture(scope="module")
def models(pkg, ng):
    out = {}
    for tag in ("shipped", "random"):
        m = pkg.network.ConvNet(channels=64, blocks=3)
        missing = m.load_state_dict(state_dict_from_golden(ng, tag))  # the reference checkpoint keys, unchanged
        m.cuda().eval()
        out[tag] = m
    return out


def test_state_dict_compat(models):
    m = models["shipped"]
    assert sum(p.numel() for p in m.parameters()) == 231820  # network.py:138
    assert len(m.state_dict()) == 50


@pytest.mark.parametrize("tag", ["shipped", "random"])
def test_forward_fp32_matches_reference_logits(models, ng, tag):
    # tolerance for the exact mode: 1e-4 absolute on log-probs (fp32 summation order only)
    lp = models[tag].forward_boards(dev_boards(ng["lg_boards"]), precision="fp32").cpu().numpy()
    ref = ng[f"lg_{tag}"]
    assert np.abs(lp - ref).max() < 1e-4
    assert (lp.argmax(1) == ref.argmax(1)).mean() >= 0.9995
    assert np.allclose(np.exp(lp).sum(1), 1.0, atol=1e-5)


def test_forward_onehot_interface(models, orc, ng):
    # the reference call: model(onehot (N,16,4,4) float32) -> (N,4) log-probs (network.py:78-87)
    b = ng["lg_boards"][:257]
    x = orc.onehot(b).cuda()
    with torch.no_grad():
        lp = models["shipped"](x)
    assert lp.shape == (257, 4) and np.abs(lp.cpu().numpy() - ng["lg_shipped"][:257]).max() < 1e-4
    models["shipped"].train()
    with pytest.raises(NotImplementedError):
        models["shipped"](x)
    models["shipped"].eval()


def test_forward_ragged_sizes(models, ng):
    m = models["random"]
    for n in (1, 7, 8, 9, 63):
        lp = m.forward_boards(dev_boards(ng["lg_boards"][:n]), precision="fp32").cpu().numpy()
        assert np.abs(lp - ng["lg_random"][:n]).max() < 1e-4


def test_policy_games_fp32_replay_reference(pkg, models, ng):
    # eval_nn-style games under identical injected spawns: exact replay of the reference's games
    seed = int(ng["pg_seed"])
    n = ng["pg_final"].size
    r = pkg.engine.playout_nn(models["shipped"].blob("fp32"), "fp32", n, seed, 0)
    same = (r["final"].cpu().numpy().view(np.uint64) == ng["pg_final"]) & \
           (r["score"].cpu().numpy().astype(np.uint64) == ng["pg_score"]) & \
           (r["moves"].cpu().numpy().astype(np.uint64) == ng["pg_moves"])
    # a game can legitimately diverge only at an fp32 near-tie (|gap| ~1e-6); allow one of 24
    assert same.sum() >= n - 1, same
    assert int(r["total_moves"].item()) == int(r["moves"].sum().item())
    g, s, c = pkg.mcts_nn.play_nn(models["shipped"], seed=seed, gid=0, precision="fp32")
    assert (g, s, c) == (int(ng["pg_final"][0]), int(ng["pg_score"][0]), int(ng["pg_moves"][0]))


def test_mcts_nn_fp32_matches_reference(pkg, models, ng):
    number, seed = int(ng["mn_number"]), int(ng["mn_seed"])
    for e, o in enumerate(ng["mn_origins"].tolist()):
        got = pkg.mcts_nn.mcts_nn(models["random"], o, number, device="cuda", seed=seed, eval_id=e, precision="fp32")
        assert got == ng["mn_result_random"][e].tolist()


def test_eval_nn_and_min(pkg, models, orc, ng):
    m = models["random"]
    mean_log, mx, mean_moves = pkg.eval_nn.eval_nn(m, number=64, seed=11, precision="fp32")
    sd = state_dict_from_golden(ng, "random")
    keys = np.array([orc.rng_key(11, g) for g in range(64)], np.uint64)
    starts = np.array([orc.init_board(int(k)) for k in keys], np.uint64)
    f, s, mv, _ = orc.nn_games(sd, starts, keys)
    assert abs(mean_log - np.mean(np.log10(s.astype(np.float64) + 1))) < 1e-9
    assert mx == s.max() and abs(mean_moves - mv.mean()) < 1e-9
    # eval_nn_min: mean over repeats of min_j L_j (eval_nn.py:76-94)
    got = pkg.eval_nn.eval_nn_min(m, number=8, repeats=4, seed=11, precision="fp32")
    assert got == np.mean([mv[k * 8:(k + 1) * 8].min() for k in range(4)])
