import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure)."""
    from oracle import oracle

    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def eg():
    return np.load(os.path.join(GOLDEN, "engine_golden.npz"))


@pytest.fixture(scope="session")
def ng():
    return np.load(os.path.join(GOLDEN, "c64b3_golden.npz"))


@pytest.fixture(scope="session")
def lg100k():
    """>= 1e5 on-distribution boards with the unmodified reference network's log-probs (oracle/make_golden_logits.py)."""
    return np.load(os.path.join(GOLDEN, "c64b3_logits_100k.npz"))


@pytest.fixture(scope="session")
def pkg():
    """The product package (directory name starts with a digit -> importlib)."""
    return importlib.import_module("2048nn_b200")


def state_dict_from_golden(ng, tag):
    import torch

    pfx = f"sd_{tag}/"
    return {k[len(pfx):]: torch.from_numpy(np.array(ng[k])) for k in ng.files if k.startswith(pfx)}
