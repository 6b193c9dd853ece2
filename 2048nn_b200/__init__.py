"""2048nn_b200 -- B200-native (sm_100a) Monte-Carlo playout engine for fqjin/2048NN.

Drop-in modules for the reference's entry points live beside this file (`board`, `mcts`,
`mcts_nn`, `eval_nn`, `network`, `selfplay`); `dropin/` holds bare-name shims so that
`from board import *` style scripts run unchanged.  The compute path is hand-written CUDA in
csrc/ behind the C ABI of include/b2048.h; there is no CPU fallback.

The directory name starts with a digit, so import it with
    importlib.import_module("2048nn_b200")
"""
from . import _lib
from ._lib import B2048Error, LIB_PATH, load_library

__all__ = ["B2048Error", "LIB_PATH", "load_library", "engine"]


def __getattr__(name):
    import importlib

    if name in ("engine", "board", "mcts", "mcts_nn", "eval_nn", "network", "selfplay", "dist", "convnet", "game_dataset", "evolve", "trainer", "analysis"):
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
