"""ctypes binding of lib2048nn_b200.so (include/b2048.h).

There is NO fallback: if the shared library is missing or no CUDA device is present, every
entry point raises.  PyTorch is used only for device memory, streams and (elsewhere)
torch.distributed.
"""
import ctypes as C
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# B2048_LIB selects another build of the same library (the -DB2048_DEBUG_ASSERTS variant of tools/debug_asserts_run.sh)
LIB_PATH = os.environ.get("B2048_LIB") or os.path.join(_HERE, "lib2048nn_b200.so")

u64, u32, i64, P = C.c_uint64, C.c_uint32, C.c_int64, C.c_void_p


class B2048Error(RuntimeError):
    pass


_lib = None
_ctx = {}

_SIGS = {
    "b2048_abi_version": (C.c_int, []),
    "b2048_last_error": (C.c_char_p, []),
    "b2048_create": (C.c_int, [C.c_int, C.POINTER(P)]),
    "b2048_destroy": (C.c_int, [P]),
    "b2048_num_sms": (C.c_int, [P]),
    "b2048_launch_count": (i64, [P]),
    "b2048_probe_pipe_rate": (C.c_int, [P, C.c_int, u32, P, C.POINTER(u64), P]),
    "b2048_set_playout_variant": (C.c_int, [P, C.c_int]),
    "b2048_set_latency_mode": (C.c_int, [P, C.c_int]),
    "b2048_take_error_flag": (C.c_int, [P, P, C.POINTER(C.c_int)]),
    "b2048_tables": (C.c_int, [P, P, P, P, P, P, P]),
    "b2048_move_batch": (C.c_int, [P, P, P, C.c_int, P, P, P, i64, P]),
    "b2048_board_ops": (C.c_int, [P, P, P, P, P, i64, P]),
    "b2048_init_boards": (C.c_int, [P, u64, u64, P, i64, P]),
    "b2048_spawn_batch": (C.c_int, [P, P, u64, P, u64, P, u32, P, P, i64, P]),
    "b2048_policy_step": (C.c_int, [P, P, P, u64, P, u64, P, P, P, P, i64, P]),
    "b2048_encode_boards": (C.c_int, [P, P, C.c_int, P, i64, P]),
    "b2048_soft_targets": (C.c_int, [P, P, C.c_float, P, i64, P]),
    "b2048_rng_draws": (C.c_int, [P, u64, u64, u64, C.c_int, P, i64, P]),
    "b2048_playout_fixed": (C.c_int, [P, P, P, u64, u64, u32, P, P, P, i64, P]),
    "b2048_convnet_blob_bytes": (i64, [C.c_int]),
    "b2048_convnet_forward": (C.c_int, [P, C.c_int, P, P, P, i64, P]),
    "b2048_convnet_tc_debug": (C.c_int, [P, P, P, P, P, C.c_int, i64, P]),
    "b2048_playout_nn": (C.c_int, [P, C.c_int, P, P, u64, u64, u32, C.c_int, P, P, P, P, P, i64, P]),
    "b2048_playout_nn_population": (C.c_int, [P, C.c_int, P, i64, C.c_int, P, u64, u64, u64, C.c_int, P, P, P, P, P, i64, P]),
    "b2048_mcts_nn": (C.c_int, [P, C.c_int, P, P, i64, C.c_int, C.c_int, C.c_int, u64, u64, P, P, P, P, P, P]),
    "b2048_selfplay_fixed": (C.c_int, [P, P, i64, u64, u64, C.c_int, C.c_int, C.c_int, u64, C.c_int, P, P, P, P, P, P, P, P]),
    "b2048_selfplay_nn_workspace_bytes": (i64, [i64, C.c_int]),
    "b2048_selfplay_nn": (C.c_int, [P, C.c_int, P, P, i64, u64, u64, C.c_int, u64, C.c_int, P, P, P, P, P, P, P, P, P]),
    "b2048_game_summary": (C.c_int, [P, P, P, P, P, P, i64, P]),
    "b2048_train_workspace_bytes": (i64, [i64]),
    "b2048_train_grad_offset": (i64, [i64]),
    "b2048_train_logp_offset": (i64, [i64]),
    "b2048_train_forward_backward": (C.c_int, [P, P, P, P, i64, P, P, i64, C.c_float, C.c_int, P, P]),
    "b2048_train_adam": (C.c_int, [P, P, P, P, P, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, i64, P]),
    "b2048_mcts_fixed": (C.c_int, [P, P, i64, C.c_int, C.c_int, C.c_int, u64, u64, P, P, P, P, P, P, P, P]),
}


def exported_symbols():
    """Names include/b2048.h declares (used by the CPU symbol test)."""
    return sorted(_SIGS)


def load_library():
    """dlopen the in-tree library and set prototypes.  Does not need a GPU."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B2048Error(
                f"{LIB_PATH} is missing: build it with `python 2048nn_b200/build.py` "
                "(or __graft_entry__.build()); there is no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)  # AttributeError = symbol missing
            fn.restype, fn.argtypes = res, args
        if lib.b2048_abi_version() != 1:
            raise B2048Error("lib2048nn_b200.so ABI version mismatch")
        _lib = lib
    return _lib


def check(rc):
    if rc:
        raise B2048Error(load_library().b2048_last_error().decode())


def require_cuda():
    if not torch.cuda.is_available():
        raise B2048Error("2048nn_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    load_library()


def ctx(device=None):
    """Per-device engine context (row tables are built on the device at creation)."""
    require_cuda()
    dev = torch.cuda.current_device() if device is None else torch.device(device).index
    if dev is None:
        dev = torch.cuda.current_device()
    if dev not in _ctx:
        lib = load_library()
        torch.cuda.init()
        h = P()
        check(lib.b2048_create(dev, C.byref(h)))
        _ctx[dev] = h
    return _ctx[dev]


def launch_count(device=None):
    """Kernels launched through the device's context so far (b2048_launch_count)."""
    return int(load_library().b2048_launch_count(ctx(device)))


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def take_error_flag(device=None):
    f = C.c_int(0)
    check(load_library().b2048_take_error_flag(ctx(device), stream_ptr(device), C.byref(f)))
    return f.value


# ---- uint64 <-> int64 bit-pattern helpers (torch has weak uint64 support) ----------------
def to_i64(x):
    """python ints / numpy uint64 -> numpy int64 with the same bits"""
    return np.asarray(x, dtype=np.uint64).view(np.int64)


def boards_to_device(x, device="cuda"):
    require_cuda()
    if isinstance(x, torch.Tensor):
        if x.dtype != torch.int64:
            raise TypeError("board tensors must be int64 bit patterns")
        return x.to(device).contiguous()
    return torch.from_numpy(np.ascontiguousarray(to_i64(x)).reshape(-1)).to(device)


def boards_to_host(t):
    return t.detach().cpu().numpy().view(np.uint64)
