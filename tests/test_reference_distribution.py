"""KS tests against samples the UNMODIFIED reference produced under its OWN RNG (CPython MT19937,
oracle/make_golden_mt.py): north-star bullet 4 -- 'under independent RNG, the score distribution must be
statistically indistinguishable from the reference's'."""
import os

import numpy as np
import pytest
from scipy.stats import ks_2samp

from conftest import state_dict_from_golden

MT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_mt_samples.npz")


def test_oracle_fixed_policy_vs_reference_mt(orc):
    g = np.load(MT)
    f, s, c = orc.play_fixed_batch(20000, 2024, 0, nthreads=4)
    assert ks_2samp(np.log10(s.astype(float) + 1), np.log10(g["fixed_score"].astype(float) + 1)).pvalue > 1e-3
    assert ks_2samp(c.astype(float), g["fixed_moves"].astype(float)).pvalue > 1e-3


@pytest.mark.gpu
def test_gpu_fixed_policy_vs_reference_mt(pkg):
    g = np.load(MT)
    f, s, c = pkg.engine.playout_fixed(200_000, 77, 0)
    gs = np.log10(s.cpu().numpy().astype(float) + 1)
    assert ks_2samp(gs, np.log10(g["fixed_score"].astype(float) + 1)).pvalue > 1e-3
    assert ks_2samp(c.cpu().numpy().astype(float), g["fixed_moves"].astype(float)).pvalue > 1e-3
    assert abs(gs.mean() - np.log10(g["fixed_score"].astype(float) + 1).mean()) < 0.02


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp16", "fp32"])
def test_gpu_nn_policy_vs_reference_mt(pkg, ng, precision):
    g = np.load(MT)
    m = pkg.network.ConvNet(64, 3)
    m.load_state_dict(state_dict_from_golden(ng, "shipped"))
    m.cuda().eval()
    n = 4000 if precision == "fp16" else 600
    r = pkg.engine.playout_nn(m.blob(precision), precision, n, 31, 0)
    s = r["score"].cpu().numpy().astype(float)
    mv = r["moves"].cpu().numpy().astype(float)
    assert ks_2samp(np.log10(s + 1), np.log10(g["nn_score"] + 1.0)).pvalue > 1e-3
    assert ks_2samp(mv, g["nn_moves"].astype(float)).pvalue > 1e-3
