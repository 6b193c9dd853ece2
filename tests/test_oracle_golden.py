"""CPU: pins the C/torch oracle to golden vectors produced by the unmodified reference
(oracle/make_golden.py) and to the known answers of SURVEY.md Appendix B."""
import hashlib

import numpy as np
import pytest

from conftest import state_dict_from_golden


def test_row_tables_match_reference(orc, eg):
    ff, fr, sf, sr, mf, mr = orc.tables()
    assert np.array_equal(ff, eg["fin_fwd"]) and np.array_equal(fr, eg["fin_rev"])
    assert np.array_equal(sf, eg["score_fwd"]) and np.array_equal(sr, eg["score_rev"])
    assert np.array_equal(mf, eg["moved_fwd"]) and np.array_equal(mr, eg["moved_rev"])


def test_row_table_known_answers(orc):
    # SURVEY.md Appendix B checksums of merge_table / merge_table_rev (board.py:138-148)
    ff, fr, sf, sr, mf, mr = orc.tables()
    assert (ff[0x1111], sf[0x1111], mf[0x1111]) == (0x0022, 8, 1)
    assert (fr[0x1111], sr[0x1111], mr[0x1111]) == (0x2200, 8, 1)
    assert (ff[0xFFFF], sf[0xFFFF], mf[0xFFFF]) == (0, 131072, 1)
    assert (ff[0x0012], sf[0x0012], mf[0x0012]) == (0x0012, 0, 0)
    assert int(ff.astype(np.int64).sum()) == 1484873157 and int(fr.astype(np.int64).sum()) == 2288378112
    assert int(sf.astype(np.int64).sum()) == 100660224 and np.array_equal(sf, sr)
    assert int(mf.sum()) == 21210 and int(mr.sum()) == 21210
    assert hashlib.sha256(ff.astype("<u2").tobytes()).hexdigest()[:16] == "2311124775fc20af"
    assert hashlib.sha256(fr.astype("<u2").tobytes()).hexdigest()[:16] == "7cce68ef7c911faf"
    assert hashlib.sha256(sf.astype("<u4").tobytes()).hexdigest()[:16] == "5998c5d66dcf2d4a"


def test_bit_tricks(orc, eg):
    assert orc.transpose(0x0123456789ABCDEF) == 0x048C159D26AE37BF
    assert orc.countzero(0) == 0 and orc.countzero(0x2141348) == 9
    b = eg["mv_boards"]
    assert [orc.transpose(x) for x in b[:3000].tolist()] == eg["transpose"][:3000].tolist()
    assert [orc.countzero(x) for x in b[:3000].tolist()] == eg["countzero"][:3000].tolist()


def test_move_vectors(orc, eg):
    b = eg["mv_boards"]
    for d in range(4):
        ob, os_, om = orc.move_batch(b, d)
        assert np.array_equal(ob, eg["mv_out"][d])
        assert np.array_equal(os_, eg["mv_score"][d])
        assert np.array_equal(om, eg["mv_moved"][d])
    # Appendix B
    assert orc.move(0x2141348, 2) == (0x21401348, 0, True)
    assert orc.move(0x2141348, 3) == (0x1214034800000000, 0, True)
    assert orc.move(0x2141348, 0) == (0x2141348, 0, False)
    assert orc.move(0xFFFFFFFFFFFFFFFF, 1) == (0, 524288, True)
    assert orc.move(0x1111222233334444, 0) == (0x0022003300440055, 120, True)
    assert orc.move(0x1111222233334444, 2) == (0x2200330044005500, 120, True)
    assert orc.move(0x0123456789ABCDEF, 3) == (0x41238567C9AB0DEF, 0, True)
    assert orc.move(0x2141348, 6)[0] == orc.move(0x2141348, 2)[0]  # direction mod 4 (board.py:161)


def test_generate_tile(orc, eg):
    # draws (pos=3, ten=1) / (pos=0, ten=0) of Appendix B expressed as words
    w = lambda v, n: ((v << 32) // n) + 1
    assert orc.generate_tile(0x2141348, w(3, 9), w(1, 10)) == 0x10002141348
    assert orc.generate_tile(0x2141348, 0, 0) == 0x22141348
    with pytest.raises(ValueError):
        orc.generate_tile(0, 0, 0)  # 16 empties wrap to 0 -> randrange(0)
    for b, k, o in zip(eg["sp_in"].tolist(), eg["sp_key"].tolist(), eg["sp_out"].tolist()):
        assert orc.spawn(b, k, 0) == o


def test_injected_play_fixed(orc, eg):
    seed = int(eg["pf_seed"])
    f, s, c = orc.play_fixed_batch(eg["pf_start"].size, seed, 0, nthreads=2)
    assert np.array_equal(f, eg["pf_final"]) and np.array_equal(s, eg["pf_score"])
    assert np.array_equal(c.astype(np.uint64), eg["pf_count"])
    starts = [orc.init_board(orc.rng_key(seed, g)) for g in range(64)]
    assert starts == eg["pf_start"][:64].tolist()
    n = eg["po_final"].size
    f, s, c = orc.play_fixed_batch(n, int(eg["po_seed"]), 0, starts=np.full(n, int(eg["po_origin"]), np.uint64))
    assert np.array_equal(f, eg["po_final"]) and np.array_equal(s, eg["po_score"])
    assert np.array_equal(c.astype(np.uint64), eg["po_count"])


def test_root_search(orc, eg):
    number, seed = int(eg["mc_number"]), int(eg["mc_seed"])
    for e, o in enumerate(eg["mc_origins"].tolist()):
        legal, rs, sc, mv, _ = orc.root_lines(o, number, seed, e)
        assert np.array_equal(sc.astype(np.int64), eg["mc_scores"][e])
        assert np.array_equal(mv.astype(np.int64), eg["mc_moves"][e])
        got = orc.mcts_fixed(o, number, seed, e)
        assert [float(x) for x in got] == eg["mc_mean"][e].tolist()  # bit-exact float64
        assert orc.mcts_fixed_min(o, number, seed, e) == eg["mc_min"][e].tolist()
        assert orc.mcts_fixed_moves(o, number, seed, e) == eg["mc_med"][e].tolist()
    o = int(eg["mc_origins"][0])
    assert [float(x) for x in orc.mcts_fixed(o, 7, 4243, 0)] == eg["mc7_mean"].tolist()
    assert orc.mcts_fixed_min(o, 7, 4243, 0) == eg["mc7_min"].tolist()
    assert orc.mcts_fixed_moves(o, 7, 4243, 0) == eg["mc7_med"].tolist()


def test_fixed_policy_statistics(orc):
    # notebooks/compare_models.ipynb:171,1463 -> mean log10(score+1) = 3.3546 / 3.3508 (5000 games);
    # distribution2.ipynb:998-999 max-tile fractions.  Independent streams -> statistical.
    f, s, c = orc.play_fixed_batch(20000, 1, 0, nthreads=4)
    m = np.log10(s.astype(np.float64) + 1).mean()
    assert abs(m - 3.3527) < 0.02
    assert 215 < c.mean() < 233
    mx = np.zeros(f.size, np.int64)
    for k in range(16):
        mx = np.maximum(mx, ((f >> np.uint64(4 * k)) & np.uint64(15)).astype(np.int64))
    frac = {t: float((mx == t).mean()) for t in (6, 7, 8, 9, 10)}
    assert abs(frac[8] - 0.4704) < 0.03 and abs(frac[7] - 0.388) < 0.03 and frac[10] < 0.002


def test_network_oracle_matches_reference_logits(orc, ng):
    import torch

    for tag in ("shipped", "random"):
        sd = state_dict_from_golden(ng, tag)
        assert sum(v.numel() for k, v in sd.items() if "running" not in k and "tracked" not in k) == 231820
        lp = orc.convnet_forward(sd, orc.onehot(ng["lg_boards"])).numpy()
        ref = ng[f"lg_{tag}"]
        assert np.abs(lp - ref).max() < 2e-5
        assert (lp.argmax(1) == ref.argmax(1)).mean() > 0.9999


def test_nn_policy_games_match_reference(orc, ng):
    sd = state_dict_from_golden(ng, "shipped")
    seed = int(ng["pg_seed"])
    k = 6
    keys = np.array([orc.rng_key(seed, g) for g in range(k)], np.uint64)
    starts = ng["pg_start"][:k]
    assert [orc.init_board(int(x)) for x in keys] == starts.tolist()
    f, s, m, _ = orc.nn_games(sd, starts, keys)
    assert np.array_equal(f, ng["pg_final"][:k]) and np.array_equal(s, ng["pg_score"][:k])
    assert np.array_equal(m.astype(np.uint64), ng["pg_moves"][:k])


def test_mcts_nn_oracle_matches_reference(orc, ng):
    sd = state_dict_from_golden(ng, "random")
    number, seed = int(ng["mn_number"]), int(ng["mn_seed"])
    for e, o in enumerate(ng["mn_origins"].tolist()):
        out = []
        for i in range(4):
            b, _, mv = orc.move(o, i)
            if not mv:
                out.append(0)
                continue
            keys = np.array([orc.rng_key(seed, orc.line_gid(e, i, j, number)) for j in range(number)], np.uint64)
            starts = np.array([orc.spawn(b, int(kk), 0) for kk in keys], np.uint64)
            _, _, m, _ = orc.nn_games(sd, starts, keys, spawn0=1)
            out.append(1 + int(m.min()))
        assert out == ng["mn_result_random"][e].tolist()
