"""GPU: record reader against the reference's own GameDataset outputs."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dataset_golden.npz")


def test_game_dataset_matches_reference(pkg, tmp_path):
    g = np.load(GOLDEN)
    for i in range(2):
        pkg.selfplay.save_record(str(tmp_path / str(i).zfill(5)), dict(boards=g[f"rec{i}_boards"], results=g[f"rec{i}_results"], score=7))
    gd = pkg.game_dataset
    path = str(tmp_path) + "/"
    for name, cls in (("plain", gd.GameDataset), ("conv", gd.ConvGameDataset), ("onehot", gd.OneHotConvGameDataset)):
        ds = cls(path, 0, 2, device="cuda", soft=3.0)
        assert tuple(ds.boards.shape) == g[f"{name}_boards"].shape and len(ds) == g[f"{name}_results"].shape[0]
        assert np.array_equal(ds.boards.cpu().numpy(), g[f"{name}_boards"])  # integer-valued floats: exact
        assert np.abs(ds.results.cpu().numpy() - g[f"{name}_results"]).max() < 2e-6  # fp32 softmax
        b, r = ds[3]
        assert b.shape == ds.boards.shape[1:] and r.shape == (4,)
    ds = gd.GameDataset(path, 1, 2, soft=1.5)
    assert np.abs(ds.results.cpu().numpy() - g["soft15_results"]).max() < 2e-6


def test_encode_matches_network_input(pkg, orc, ng):
    # the one-hot planes are exactly what mcts_nn.py:25-35 feeds the network
    b = ng["lg_boards"][:1000]
    x = pkg.engine.encode_boards(torch.from_numpy(b.view(np.int64).copy()).cuda(), onehot=True)
    assert torch.equal(x.cpu(), orc.onehot(b))
