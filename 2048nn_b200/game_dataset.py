"""Drop-in for the reference's `game_dataset` module (game_dataset.py:6-48): self-play records
(`NNNNN.npz` with boards uint64[T], results float32[T,4], score) -> training tensors on the B200.

The npz files are read on the host (as the reference does); nibble unpacking / one-hot encoding and the
soft targets run as CUDA kernels (b2048_encode_boards, b2048_soft_targets) on the concatenated record.
This is the reader side of the record format `selfplay.py` writes."""
import numpy as np
import torch
from torch.utils.data import Dataset

from . import _lib as L
from . import engine


class GameDataset(Dataset):
    """boards: float32 (N,16) tile exponents; results: float32 (N,4) soft targets."""
    _onehot = False

    def __init__(self, gamepath, start, end, device='cuda', soft=3.0):
        boards, results = [], []
        for i in range(start, end):
            game = np.load(gamepath + str(i).zfill(5) + '.npz')
            boards.append(game['boards'])  # uint64
            results.append(np.asarray(game['results'], dtype=np.float32).reshape(-1, 4))
        b = np.ascontiguousarray(np.concatenate(boards).astype(np.uint64))
        r = np.ascontiguousarray(np.concatenate(results))
        d_boards = L.boards_to_device(b)
        self.packed = d_boards  # uint64 bit patterns (int64 tensor): what the fused training step consumes
        self.boards = engine.encode_boards(d_boards, onehot=self._onehot)
        self.results = engine.soft_targets(torch.from_numpy(r).cuda(), soft)
        if torch.device(device).type != 'cuda':
            self.boards, self.results = self.boards.to(device), self.results.to(device)

    def __len__(self):
        return self.results.size(0)

    def __getitem__(self, index):
        return self.boards[index], self.results[index]


class ConvGameDataset(GameDataset):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.boards = self.boards.view(-1, 1, 4, 4)


class OneHotConvGameDataset(GameDataset):
    _onehot = True
