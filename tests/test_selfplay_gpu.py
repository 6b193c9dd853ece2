"""GPU: batched self-play (BASELINE config 4 shape) against the same main loop run on the oracle,
and the reference's npz record format (selfplay.py:55-58)."""
import numpy as np
import pytest
import torch

from conftest import state_dict_from_golden

pytestmark = pytest.mark.gpu


def oracle_selfplay_fixed(orc, G, number, seed):
    recs = []
    boards = [orc.init_board(orc.rng_key(seed, g)) for g in range(G)]
    for g in range(G):
        b, score, k, hb, hr = boards[g], 0, 0, [], []
        step = 0
        while True:
            res = orc.mcts_fixed_min(b, number, seed, (step + 1) * G + g)
            i = int(np.argmax(res))
            f, s, m = orc.move(b, i)
            if not m:
                break
            hb.append(b)
            hr.append(res)
            b = orc.spawn(f, orc.rng_key(seed, g), k)
            score += s
            k += 1
            step += 1
        recs.append((np.array(hb, np.uint64), np.array(hr, np.float32), score))
    return recs


def test_selfplay_batch_fixed_matches_oracle(pkg, orc, tmp_path):
    G, number, seed = 3, 3, 808
    recs, _ = pkg.selfplay.selfplay_batch(G, number, None, 1, seed)
    want = oracle_selfplay_fixed(orc, G, number, seed)
    for r, (hb, hr, sc) in zip(recs, want):
        assert np.array_equal(r["boards"], hb) and np.array_equal(r["results"], hr) and r["score"] == sc
        assert len(hb) > 30
    p = str(tmp_path / "00001")
    pkg.selfplay.save_record(p, recs[0])
    z = np.load(p + ".npz")
    assert sorted(z.files) == ["boards", "results", "score"]
    assert z["boards"].dtype == np.uint64 and z["results"].dtype == np.float32 and z["results"].shape[1] == 4


@pytest.fixture(scope="module")
def nets(pkg, ng):
    out = {}
    for tag in ("shipped", "random"):
        m = pkg.network.ConvNet(64, 3)
        m.load_state_dict(state_dict_from_golden(ng, tag))
        out[tag] = m.cuda().eval()
    return out


def _same_records(a, b):
    return all(np.array_equal(x["boards"], y["boards"]) and np.array_equal(x["results"], y["results"])
               and x["score"] == y["score"] for x, y in zip(a, b)) and len(a) == len(b)


def test_selfplay_nn_device_matches_oracle_loop(pkg, orc, ng, nets):
    # config 4's record shape in exact (fp32) mode, move for move against the oracle's selfplay.py:62-115 loop
    # (torch fp32 network, per-line injected streams).  Bounded to the first main moves: a root evaluation of the
    # shipped net costs the CPU oracle ~1 s.
    G, number, seed, steps = 2, 3, 606, 6
    sd = state_dict_from_golden(ng, "shipped")
    want = orc.selfplay_nn(sd, G, number, seed, steps)
    blob = nets["shipped"].blob("fp32")
    r = pkg.engine.selfplay_nn_games(blob, "fp32", G, number, seed, max_steps=steps)
    torch.cuda.synchronize()
    assert pkg._lib.take_error_flag() == 0
    for g, (hb, hr, sc, fin) in enumerate(want):
        n = int(r["steps"][g])
        assert n == len(hb) == steps and int(r["status"][g]) == 1  # stopped by the step budget, like the oracle run
        assert np.array_equal(r["boards"][g, :n].cpu().numpy().view(np.uint64), hb)
        assert np.array_equal(r["results"][g, :n].cpu().numpy(), hr)
        assert int(r["score"][g]) == sc and int(pkg.board._host(r["final"][g:g + 1])[0]) == fin
    assert int(r["total_moves"].item()) > 100


@pytest.mark.parametrize("precision,tag,G,number", [("fp32", "random", 4, 5), ("fp16", "random", 5, 6), ("fp16", "random", 70, 40)])
def test_selfplay_nn_device_equals_lockstep_loop(pkg, nets, precision, tag, G, number):
    # whole games: the device-resident queue scheduler plays exactly the games of the step-by-step host loop
    # (b2048_mcts_nn + move + spawn per main move), in exact mode and on the tensor cores; the last case has more
    # lines in flight (11,200) than one GPU has game slots (9,472), i.e. the queue is really queueing
    m = nets[tag]
    cap = 100000 if G < 10 else 12
    dev, mv_dev = pkg.selfplay.selfplay_batch(G, number, m, seed=17, precision=precision, max_steps=cap)
    host, mv_host = pkg.selfplay.selfplay_batch(G, number, m, seed=17, precision=precision, max_steps=cap, lockstep=True)
    assert _same_records(dev, host)
    assert all(len(r["boards"]) > 10 for r in dev) and mv_dev > 1000
    assert all(r["truncated"] == (G >= 10) for r in dev)
    # the early stop prunes lines differently under different schedules, the minima (and so the games) are the same
    assert mv_dev > 0 and mv_host > 0


def test_selfplay_nn_shards_of_a_session_play_the_same_games(pkg, nets):
    blob = nets["random"].blob("fp16")
    whole = pkg.engine.selfplay_nn_games(blob, "fp16", 6, 5, 4, max_steps=4096)
    part = pkg.engine.selfplay_nn_games(blob, "fp16", 3, 5, 4, game_id_base=3, games_total=6, max_steps=4096)
    assert torch.equal(whole["final"][3:], part["final"]) and torch.equal(whole["score"][3:], part["score"])
    assert torch.equal(whole["steps"][3:], part["steps"]) and int(whole["status"].sum()) == 0
    n = int(part["steps"][0])
    assert torch.equal(whole["boards"][3, :n], part["boards"][0, :n])
    assert torch.equal(whole["results"][3, :n], part["results"][0, :n])
    # a game that cannot move at all ends at once with an empty record (selfplay.py:95-98)
    dead = torch.from_numpy(np.array([0x1212212112122121], np.uint64).view(np.int64)).cuda()
    r = pkg.engine.selfplay_nn_games(blob, "fp16", 1, 5, 4, starts=dead, max_steps=16)
    assert int(r["steps"][0]) == 0 and int(r["status"][0]) == 0 and int(r["final"][0]) == int(dead[0])


def test_selfplay_record_truncation_is_flagged(pkg, nets, tmp_path):
    recs, _ = pkg.selfplay.selfplay_batch(2, 4, nets["random"], seed=3, precision="fp16", max_steps=5)
    assert all(r["truncated"] and len(r["boards"]) == 5 for r in recs)
    with pytest.raises(ValueError):
        pkg.selfplay.save_record(str(tmp_path / "x"), recs[0])
    recs, _ = pkg.selfplay.selfplay_batch(2, 3, None, 1, seed=3, max_steps=7)
    assert all(r["truncated"] for r in recs)


def test_selfplay_cli_record_is_the_same_game_alone_or_in_a_batch(pkg, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    pkg.selfplay.main(["--seed", "41", "--games", "3", "--number", "3"])
    batch = np.load("selfplay/fixed10/00042.npz")
    b42, r42, s42 = batch["boards"].copy(), batch["results"].copy(), int(batch["score"])
    pkg.selfplay.main(["--seed", "42", "--number", "3"])
    alone = np.load("selfplay/fixed10/00042.npz")
    assert np.array_equal(alone["boards"], b42) and np.array_equal(alone["results"], r42) and int(alone["score"]) == s42
