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


def test_selfplay_batch_nn_runs(pkg, ng):
    m = pkg.network.ConvNet(64, 3)
    m.load_state_dict(state_dict_from_golden(ng, "random"))
    m.cuda().eval()
    recs, total = pkg.selfplay.selfplay_batch(4, 8, m, seed=3, precision="fp16")
    assert total > 1000 and all(len(r["boards"]) > 20 for r in recs)
    for r in recs:
        assert r["results"].shape == (len(r["boards"]), 4) and (r["results"].max(1) >= 1).all()
