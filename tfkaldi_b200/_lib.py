"""ctypes binding of libtfkaldi_b200.so (the C-ABI declared in include/tfkaldi_b200.h).

There is deliberately NO fallback: if the CUDA extension is missing the import fails loudly, and if it
is present but no B200 is visible every compute entry point raises (tfk_create returns TFK_ECUDA).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtfkaldi_b200.so")

TFK_OK = 0
TFK_EINVAL, TFK_ECUDA, TFK_ENCCL, TFK_ESHAPE = -1, -2, -3, -4
TFK_ABI_VERSION = 1
TFK_PREC_BF16, TFK_PREC_BF16X3 = 0, 1
TFK_NONLIN_RELU, TFK_NONLIN_LINEAR, TFK_NONLIN_SIGMOID, TFK_NONLIN_TANH = 0, 1, 2, 3

(T_WEIGHTS, T_BIASES, T_BN_BETA, T_BN_MOVING_MEAN, T_BN_MOVING_VAR, T_ADAM_M_W, T_ADAM_V_W, T_ADAM_M_B,
 T_ADAM_V_B, T_ADAM_M_BETA, T_ADAM_V_BETA, T_GRAD_W, T_GRAD_B, T_GRAD_BETA) = range(14)
S_GLOBAL_STEP, S_LR_FACT, S_ACTIVE_LAYERS, S_LOSS_SUM, S_NUM_FRAMES, S_ADAM_STEP = range(6)
TIMER_NAMES = ["gemm_fwd", "gemm_bwd", "softmax_ce", "adam", "colsum", "bn", "convert", "decode_out", "allreduce"]
NUM_TIMERS = len(TIMER_NAMES)


class TfkConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("num_layers", C.c_int32),
        ("input_dim", C.c_int32),
        ("hidden_dim", C.c_int32),
        ("output_dim", C.c_int32),
        ("max_frames", C.c_int32),
        ("nonlin", C.c_int32),
        ("batch_norm", C.c_int32),
        ("keep_prob", C.c_float),
        ("bn_eps", C.c_float),
        ("bn_decay", C.c_float),
        ("adam_beta1", C.c_float),
        ("adam_beta2", C.c_float),
        ("adam_eps", C.c_float),
        ("precision", C.c_int32),
        ("device", C.c_int32),
        ("seed", C.c_uint64),
        ("l2_norm", C.c_int32),
        ("reserved", C.c_int32),
    ]


# every symbol include/tfkaldi_b200.h declares: (name, restype, argtypes)
_H = C.c_void_p
_FP = C.c_void_p  # raw device / host pointers are passed as integers
SIGNATURES = [
    ("tfk_default_config", None, [C.POINTER(TfkConfig)]),
    ("tfk_create", C.c_int, [C.POINTER(TfkConfig), C.POINTER(_H)]),
    ("tfk_destroy", C.c_int, [_H]),
    ("tfk_last_error", C.c_char_p, [_H]),
    ("tfk_set_tensor", C.c_int, [_H, C.c_int, C.c_int, _FP, C.c_size_t, C.c_void_p]),
    ("tfk_get_tensor", C.c_int, [_H, C.c_int, C.c_int, _FP, C.c_size_t, C.c_void_p]),
    ("tfk_set_scalar", C.c_int, [_H, C.c_int, C.c_double]),
    ("tfk_get_scalar", C.c_int, [_H, C.c_int, C.POINTER(C.c_double), C.c_void_p]),
    ("tfk_fflayer_fwd", C.c_int, [_H, C.c_int, _FP, _FP, C.c_int, C.c_int, C.c_void_p]),
    ("tfk_fflayer_bwd", C.c_int, [_H, C.c_int, _FP, _FP, C.c_int, C.c_void_p]),
    ("tfk_softmax_ce", C.c_int, [_H, _FP, _FP, C.c_int, _FP, _FP, C.c_void_p]),
    ("tfk_accumulate", C.c_int, [_H, _FP, _FP, C.c_int, C.c_void_p]),
    ("tfk_accumulate_raw", C.c_int, [_H, _FP, _FP, C.c_int, _FP, _FP, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    ("tfk_forward_loglik_raw", C.c_int, [_H, _FP, _FP, C.c_int, _FP, C.c_int, C.c_int, C.c_int, _FP, _FP, C.c_void_p]),
    ("tfk_forward_loglik_raw_rows", C.c_int, [_H, _FP, _FP, C.c_int, _FP, C.c_int, C.c_int, C.c_int, C.c_int, _FP, _FP, C.c_void_p]),
    ("tfk_apply", C.c_int, [_H, C.c_float, C.POINTER(C.c_float), C.c_void_p]),
    ("tfk_last_loss", C.c_int, [_H, C.POINTER(C.c_float), C.c_void_p]),
    ("tfk_train_step", C.c_int, [_H, _FP, _FP, C.c_int, C.c_float, C.POINTER(C.c_float), C.c_void_p]),
    ("tfk_train_step_raw", C.c_int, [_H, _FP, _FP, C.c_int, _FP, _FP, C.c_int, C.c_int, C.c_int, C.c_float, C.POINTER(C.c_float), C.c_void_p]),
    ("tfk_eval_accumulate", C.c_int, [_H, _FP, _FP, C.c_int, C.c_void_p]),
    ("tfk_eval_finish", C.c_int, [_H, C.POINTER(C.c_float), C.c_void_p]),
    ("tfk_forward_posteriors", C.c_int, [_H, _FP, C.c_int, _FP, C.c_void_p]),
    ("tfk_forward_loglik", C.c_int, [_H, _FP, C.c_int, _FP, _FP, C.c_void_p]),
    ("tfk_halve_lr", C.c_int, [_H]),
    ("tfk_set_active_layers", C.c_int, [_H, C.c_int]),
    ("tfk_set_dropout_seed", C.c_int, [_H, C.c_uint64]),
    ("tfk_get_activation", C.c_int, [_H, C.c_int, _FP, C.c_int, C.c_void_p]),
    ("tfk_comm_unique_id", C.c_int, [C.POINTER(C.c_uint8)]),
    ("tfk_comm_init", C.c_int, [_H, C.POINTER(C.c_uint8), C.c_int, C.c_int]),
    ("tfk_set_comm", C.c_int, [_H, C.c_void_p, C.c_int, C.c_int]),
    ("tfk_ipc_export", C.c_int, [_H, C.POINTER(C.c_uint8)]),
    ("tfk_ipc_import", C.c_int, [_H, C.POINTER(C.c_uint8), C.c_int]),
    ("tfk_enable_timers", C.c_int, [_H, C.c_int]),
    ("tfk_get_timers", C.c_int, [_H, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    ("tfk_kernel_launches", C.c_int64, [_H]),
    ("tfk_abi_version", C.c_int, []),
    ("tfk_device_count", C.c_int, []),
]

_lib = None


def load():
    """Load the shared library (once).  Raises ImportError with build instructions if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the sm_100a extension first "
            "(`python -c 'import __graft_entry__ as g; g.build()'` or `make -C tfkaldi_b200/csrc`). "
            "tfkaldi_b200 has no CPU or PyTorch fallback path."
        )
    lib = C.CDLL(LIB_PATH)
    for name, restype, argtypes in SIGNATURES:
        fn = getattr(lib, name)  # AttributeError if the ABI and the header drift apart
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.tfk_abi_version() != TFK_ABI_VERSION:
        raise ImportError(f"ABI mismatch: library {lib.tfk_abi_version()} vs binding {TFK_ABI_VERSION}")
    _lib = lib
    return lib


class TfkError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"tfkaldi_b200 error {code}: {msg}")
        self.code = code
