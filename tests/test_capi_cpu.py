"""CPU: the C-ABI library loads and exports every symbol include/b2048.h declares; the product
fails loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as ge

    ge.build()
    return True


def header_symbols():
    src = open(os.path.join(ROOT, "include", "b2048.h")).read()
    return sorted(set(re.findall(r"\b(b2048_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built, pkg):
    lib = ctypes.CDLL(pkg.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), s
    assert lib.b2048_abi_version() == 1
    # the Python binding table covers the header too
    assert sorted(pkg._lib.exported_symbols()) == syms


def test_sass_is_sm100a(built, pkg):
    import subprocess

    out = subprocess.run(["cuobjdump", "-lelf", pkg.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu(built, pkg):
    with pytest.raises(pkg.B2048Error):
        pkg.board.move(0x2141348, 0)
    with pytest.raises(pkg.B2048Error):
        pkg.board.play_fixed_batch(4)


def test_product_never_imports_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "2048nn_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle|oracle/|libb2048_oracle", txt, re.M):
                    bad.append(f)
    assert not bad, bad


def test_header_is_plain_c(tmp_path):
    # the boundary is a C ABI: the header must compile as C99 with no C++ / torch types
    import subprocess

    src = tmp_path / "hdr.c"
    src.write_text('#include "b2048.h"\nint main(void) { return b2048_abi_version() == 0; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_new_entry_points_fail_loudly_without_gpu(built, pkg):
    # trainer, population evaluation, analysis and device-resident games: no CPU fallback either
    if torch.cuda.is_available():
        pytest.skip("checks the no-GPU failure mode")
    m = pkg.network.ConvNet(64, 3)
    with pytest.raises(pkg.B2048Error):
        pkg.trainer.TrainStep(m)
    with pytest.raises(pkg.B2048Error):
        pkg.mcts.play_mcts_fixed(number=2, seed=1)
    with pytest.raises(pkg.B2048Error):
        pkg.analysis.fixed_distribution(8, seed=1)
    with pytest.raises(pkg.B2048Error):
        pkg.evolve.eval_population([m], number=2, seed=1)


def test_trainer_host_logic(pkg):
    # the parts of the training driver that need no GPU: run name grammar (trainer.py:17-21,38), the patience rule
    # (trainer.py:46-52,66-67,81-91) and the up-front architecture check
    tr = pkg.trainer
    args = tr.parser().parse_args(["--name", "x", "--t_tuple", "0", "3400", "--seed", "0"])
    assert tr.run_name(args, False) == "x_0_3400_soft3.0c64b3_p10_bs2048lr0.1d0.0_s0"
    assert tr.run_name(tr.parser().parse_args(["--t_tuple", "0", "3400"]), True) == "pre_0_3400_soft3.0c64b3_p10_bs2048lr0.1d0.0_s0"
    p = tr.Patience(2, 100)
    p.after_train(0, 0.5)
    assert p.after_eval(3.0) and not p.exhausted and p.remaining == 2
    p.after_train(1, 0.4)
    assert not p.after_eval(2.0) and p.remaining == 1
    p.after_train(2, 0.3)            # third epoch still above 0.210: give up (trainer.py:66-67)
    assert p.exhausted
    q = tr.Patience(0, 7)            # patience 0 = run all epochs
    for e in range(6):
        q.after_train(e, 0.1)
        q.after_eval(0.0 if e else 1.0)
    assert not q.exhausted and q.remaining == 2
    with pytest.raises(SystemExit):
        tr.main(tr.parser().parse_args(["--channels", "16", "--blocks", "5"]))
    # the self-play command line plays main game N of one fixed session (same record alone or in a batch)
    assert pkg.selfplay.CLI_GAMES_TOTAL == 1 << 32


def test_main_game_workspace_sizes(built, pkg):
    lib = pkg._lib.load_library()
    small, big = lib.b2048_selfplay_nn_workspace_bytes(1, 50), lib.b2048_selfplay_nn_workspace_bytes(256, 50)
    assert 0 < small < big < 4 << 20 and small % 4 == 0
    assert lib.b2048_selfplay_nn_workspace_bytes(0, 50) == -1 and lib.b2048_selfplay_nn_workspace_bytes(1 << 26, 50) == -1
