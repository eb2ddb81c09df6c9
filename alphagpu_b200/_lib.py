"""ctypes binding of libalphagpu.so (include/alphagpu.h) — the Python twin of the Julia `ccall` glue.

There is no fallback: if the library cannot be loaded or no CUDA device is present, creating a
context raises.  Nothing here imports the CPU oracle.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("AGPU_LIB", os.path.join(HERE, "libalphagpu.so"))   # AGPU_LIB: development variants

OK, ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_STATE, ERR_ILLEGAL_MOVE = 0, -1, -2, -3, -4, -5
CONNECT4, GOBANG, HEX, REVERSI8, REVERSI6 = 0, 1, 2, 3, 4
NN_BF16_TC, NN_FP32, NN_FP16_TC = 0, 1, 2
NKERNELS = 8
KERNEL_CLASSES = ("select", "nn", "expand_backup", "begin", "finish_ply", "compact", "finalize", "ply_fused")


class Config(C.Structure):
    _fields_ = [("game", C.c_int32), ("n", C.c_int32), ("nvict", C.c_int32), ("rollouts", C.c_int32), ("max_games", C.c_int64),
                ("width", C.c_int32), ("blocks", C.c_int32), ("device", C.c_int32), ("nn_mode", C.c_int32)]


class GameInfo(C.Structure):
    _fields_ = [("max_actions", C.c_int32), ("vectorized_state", C.c_int32), ("feature_size", C.c_int32),
                ("max_length_game", C.c_int32), ("position_bytes", C.c_int32)]


class TreeDump(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("nnodes", "parent", "action", "child", "order", "nchild", "expanded", "prior", "q", "visits", "states")]


class Samples(C.Structure):
    _fields_ = [("capacity", C.c_int64), ("count", C.c_int64), ("state", C.c_void_p), ("policy", C.c_void_p), ("player", C.c_void_p),
                ("value", C.c_void_p), ("fstate", C.c_void_p), ("game", C.c_void_p), ("ply", C.c_void_p)]


class RunStats(C.Structure):
    _fields_ = [("sims", C.c_int64), ("positions", C.c_int64), ("plies", C.c_int64), ("total_length", C.c_int64), ("faults", C.c_int64),
                ("kernel_launches", C.c_int64), ("device_ms", C.c_double), ("search_ms", C.c_double)]


class KernelTimes(C.Structure):
    _fields_ = [("launches", C.c_int64 * NKERNELS), ("ms", C.c_double * NKERNELS), ("nodes_traversed", C.c_int64), ("descents", C.c_int64)]


class TrainConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("inp", C.c_int32), ("width", C.c_int32), ("blocks", C.c_int32), ("actions", C.c_int32),
                ("fsize", C.c_int32), ("max_batch", C.c_int32), ("reserved", C.c_int32), ("lr", C.c_double), ("beta1", C.c_double),
                ("beta2", C.c_double), ("eps", C.c_double), ("weight_decay", C.c_double), ("feature_weight", C.c_float),
                ("reserved2", C.c_float)]


# every symbol include/alphagpu.h declares, with its signature
_VP, _I32, _I64, _U32, _U64, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_float
SIGNATURES = {
    "agpu_abi_version": (C.c_int, []),
    "agpu_game_info_get": (C.c_int, [_I32, _I32, _I32, C.POINTER(GameInfo)]),
    "agpu_create": (C.c_int, [C.POINTER(_VP), C.POINTER(Config)]),
    "agpu_destroy": (None, [_VP]),
    "agpu_last_error": (C.c_char_p, [_VP]),
    "agpu_set_weights": (C.c_int, [_VP, _I32, _VP, _VP, _VP, _VP, _VP, _VP]),
    "agpu_forward": (C.c_int, [_VP, _I32, _VP, _I64, _VP, _VP]),
    "agpu_position_init": (C.c_int, [_VP, _VP, _I64]),
    "agpu_can_play": (C.c_int, [_VP, _VP, _I64, _VP]),
    "agpu_play": (C.c_int, [_VP, _VP, _VP, _I64, _VP]),
    "agpu_is_over": (C.c_int, [_VP, _VP, _I64, _VP, _VP]),
    "agpu_encode": (C.c_int, [_VP, _VP, _I64, _VP]),
    "agpu_reinit": (C.c_int, [_VP, _VP, _I64, _VP]),
    "agpu_search": (C.c_int, [_VP, _I64, _I32, _I32, _I32, _F, _F, _VP, _U64, _U32]),
    "agpu_get_roots": (C.c_int, [_VP, _I64, _VP, _VP]),
    "agpu_search_begin": (C.c_int, [_VP, _I64]),
    "agpu_select": (C.c_int, [_VP, _I64, _I32, _I32, _F, _VP, _U64, _U32]),
    "agpu_get_leaves": (C.c_int, [_VP, _I64, _VP, _VP]),
    "agpu_eval": (C.c_int, [_VP, _I64, _I32, _VP, _VP]),
    "agpu_expand_backup": (C.c_int, [_VP, _I64, _I32, _I32, _VP, _VP]),
    "agpu_get_tree": (C.c_int, [_VP, _I64, C.POINTER(TreeDump)]),
    "agpu_selfplay": (C.c_int, [_VP, _I32, _I32, _I64, _U32, _F, _F, _U64, C.POINTER(Samples), _VP, C.POINTER(RunStats)]),
    "agpu_duel": (C.c_int, [_VP, _I32, _I32, _I32, _I64, _U32, _F, _U64, _VP, C.POINTER(RunStats)]),
    "agpu_multi_create": (C.c_int, [C.POINTER(_VP), C.POINTER(Config), _I32, _VP]),
    "agpu_multi_destroy": (None, [_VP]),
    "agpu_multi_last_error": (C.c_char_p, [_VP]),
    "agpu_multi_ngpus": (C.c_int, [_VP]),
    "agpu_multi_context": (_VP, [_VP, _I32]),
    "agpu_multi_set_weights": (C.c_int, [_VP, _I32, _VP, _VP, _VP, _VP, _VP, _VP]),
    "agpu_multi_selfplay": (C.c_int, [_VP, _I32, _I32, _I64, _U32, _F, _F, _U64, C.POINTER(Samples), _VP, C.POINTER(RunStats)]),
    "agpu_multi_duel": (C.c_int, [_VP, _I32, _I32, _I32, _I64, _U32, _F, _U64, _VP, C.POINTER(RunStats)]),
    "agpu_profile": (C.c_int, [_VP, _I32]),
    "agpu_get_kernel_times": (C.c_int, [_VP, C.POINTER(KernelTimes), _I32]),
    "agpu_layout_info": (C.c_int, [_VP, C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I64)]),
    "agpu_host_alloc": (C.c_int, [C.POINTER(_VP), _U64]),
    "agpu_host_free": (C.c_int, [_VP]),
    "agpu_debug_expf": (C.c_int, [_VP, _VP, _I64, _VP, _I32]),
    "agpu_debug_fdiv_check": (C.c_int, [_U64, _U64, _VP]),
}

_lib = None


class AlphaGPUError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libalphagpu error {code}: {msg}")
        self.code = code


# include/alphagpu_train.h
_W8 = [C.c_void_p] * 8
SIGNATURES.update({
    "agpu_trainer_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(TrainConfig)]),
    "agpu_trainer_destroy": (None, [C.c_void_p]),
    "agpu_trainer_last_error": (C.c_char_p, [C.c_void_p]),
    "agpu_trainer_set_params": (C.c_int, [C.c_void_p, *_W8, C.c_int32]),
    "agpu_trainer_get_params": (C.c_int, [C.c_void_p, *_W8]),
    "agpu_trainer_get_grads": (C.c_int, [C.c_void_p, *_W8]),
    "agpu_trainer_loss_grad": (C.c_int, [C.c_void_p] * 5 + [C.c_int64, C.c_void_p]),
    "agpu_trainer_loss": (C.c_int, [C.c_void_p] * 5 + [C.c_int64, C.c_void_p]),
    "agpu_trainer_step": (C.c_int, [C.c_void_p] * 5 + [C.c_int64, C.c_void_p]),
    "agpu_trainer_grad_buffer": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "agpu_trainer_apply": (C.c_int, [C.c_void_p, C.c_float]),
    "agpu_trainer_opt_state": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "agpu_trainer_last_ms": (C.c_int, [C.c_void_p, C.c_void_p]),
})


def load():
    """dlopen libalphagpu.so (built by `python -m alphagpu_b200.build`); raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(f"{SO_PATH} not built: run `python -m alphagpu_b200.build` (there is no CPU fallback)")
        lib = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(ctx, rc, allow=()):
    if rc != OK and rc not in allow:
        msg = load().agpu_last_error(ctx)
        raise AlphaGPUError(rc, msg.decode() if msg else "")
    return rc
