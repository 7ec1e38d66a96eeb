"""ctypes binding of libqsv.so (the C ABI declared in include/qsv.h).

There is NO fallback: if the shared library is missing or no CUDA device is visible the
product path raises — it never routes through NumPy or the oracle.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_CSRC = Path(__file__).resolve().parent / "csrc"
LIB_PATH = _CSRC / os.environ.get("QSV_LIB_NAME", "libqsv.so")   # env override: A/B builds in experiments

QSV_C64, QSV_C128 = 0, 1
QSV_OK, QSV_EINVAL, QSV_ENONLOCAL, QSV_ECUDA, QSV_ENOMEM, QSV_ECOMM, QSV_EIO = 0, -1, -2, -3, -4, -5, -6
QSV_MAX_TILE_BITS, QSV_REG_BITS, QSV_MAX_ROUNDS, QSV_MAX_ACTIVE_BITS = 14, 4, 16, 52
OP_HAD, OP_ROT, OP_XSWAP, OP_YSWAP, OP_PHASE, OP_SIGN, OP_SCALE, OP_TPHASE = range(8)
OP_WITH_TARGET = (OP_HAD, OP_ROT, OP_XSWAP, OP_YSWAP)
OPF_PRESIGN, OPF_PRENEG, OPF_PREPHASE = 1, 2, 4
OPT_JIT, OPT_SIMPLE_PASS, OPT_PEER_SWAP = 1, 2, 3


class QsvOp(C.Structure):
    _fields_ = [("kind", C.c_uint8), ("target", C.c_uint8), ("reg_ctrl", C.c_uint8), ("flags", C.c_uint8),
                ("tile_ctrl", C.c_uint32), ("glob_ctrl", C.c_uint64), ("m", C.c_double * 4)]


class QsvRound(C.Structure):
    _fields_ = [("reg_pos", C.c_uint8 * QSV_REG_BITS), ("thr_pos", C.c_uint8 * QSV_MAX_TILE_BITS),
                ("op_begin", C.c_int32), ("op_end", C.c_int32), ("fold_off", C.c_int32)]


class QsvPass(C.Structure):
    _fields_ = [("n_tile", C.c_int32), ("load_bits", C.c_int32 * QSV_MAX_TILE_BITS),
                ("store_bits", C.c_int32 * QSV_MAX_TILE_BITS), ("n_rounds", C.c_int32),
                ("rounds", QsvRound * QSV_MAX_ROUNDS), ("n_ops", C.c_int32), ("n_fold", C.c_int32),
                ("n_active", C.c_int32), ("active_bits", C.c_int32 * QSV_MAX_ACTIVE_BITS), ("zero_input", C.c_int32),
                ("store_flip", C.c_uint64)]


class QsvTiming(C.Structure):
    _fields_ = [("ms", C.c_float), ("kind", C.c_int32), ("pass_index", C.c_int32)]


class QsvError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libqsv error {code}: {msg}")
        self.code = code


_H = C.c_void_p
_P = C.c_void_p
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)

# name -> (restype, argtypes); every symbol include/qsv.h declares
SIGNATURES = {
    "qsv_abi_version": (C.c_int, []),
    "qsv_device_count": (C.c_int, []),
    "qsv_create": (C.c_int, [C.POINTER(_H), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "qsv_destroy": (C.c_int, [_H]),
    "qsv_release_cached": (C.c_int, []),
    "qsv_last_error": (C.c_char_p, [_H]),
    "qsv_sync": (C.c_int, [_H]),
    "qsv_device_ptr": (C.c_int, [_H, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_void_p)]),
    "qsv_init_zero": (C.c_int, [_H]),
    "qsv_init_basis": (C.c_int, [_H, C.c_uint64]),
    "qsv_upload": (C.c_int, [_H, C.c_void_p, C.c_size_t, C.c_size_t]),
    "qsv_download": (C.c_int, [_H, C.c_void_p, C.c_size_t, C.c_size_t]),
    "qsv_upload_async": (C.c_int, [_H, C.c_void_p, C.c_size_t, C.c_size_t]),
    "qsv_download_async": (C.c_int, [_H, C.c_void_p, C.c_size_t, C.c_size_t]),
    "qsv_snapshot": (C.c_int, [_H]),
    "qsv_snapshot_download_async": (C.c_int, [_H, C.c_void_p, C.c_size_t, C.c_size_t]),
    "qsv_snapshot_sync": (C.c_int, [_H]),
    "qsv_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "qsv_host_free": (C.c_int, [C.c_void_p]),
    "qsv_apply_1q": (C.c_int, [_H, C.c_int, _dp]),
    "qsv_apply_2q": (C.c_int, [_H, C.c_int, C.c_int, _dp]),
    "qsv_apply_diag": (C.c_int, [_H, C.c_int, _ip, _dp]),
    "qsv_apply_ctrl_1q": (C.c_int, [_H, C.c_int, C.c_int, _dp]),
    "qsv_apply_kq": (C.c_int, [_H, C.c_int, _ip, _dp]),
    "qsv_apply_pass": (C.c_int, [_H, C.POINTER(QsvPass), C.POINTER(QsvOp), _dp]),
    "qsv_program_create": (C.c_int, [_H, C.POINTER(QsvPass), C.c_int, C.POINTER(QsvOp), _dp, C.POINTER(_P)]),
    "qsv_program_run": (C.c_int, [_H, _P]),
    "qsv_program_destroy": (C.c_int, [_H, _P]),
    "qsv_program_run_range": (C.c_int, [_H, _P, C.c_int, C.c_int]),
    "qsv_program_specialised": (C.c_int, [_H, _P, _ip, C.c_int]),
    "qsv_swap_pipelined": (C.c_int, [_H, _P, C.c_int, C.c_int, _P, C.c_int, C.c_int, _ip, _ip, C.c_int, _ip, C.c_int]),
    "qsv_shadow_ptr": (C.c_int, [_H, C.POINTER(C.c_void_p)]),
    "qsv_comm_shadow_ipc_handle": (C.c_int, [_H, C.c_void_p]),
    "qsv_comm_set_shadow_peers": (C.c_int, [_H, C.c_void_p]),
    "qsv_scatter_set_targets": (C.c_int, [_H, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "qsv_pass_scatter_prepare": (C.c_int, [_H, _P, C.c_int, C.c_int, _ip]),
    "qsv_pass_scatter": (C.c_int, [_H, _P, C.c_int, C.c_int, _ip, _ip, _ip]),
    "qsv_set_option": (C.c_int, [_H, C.c_int, C.c_longlong]),
    "qsv_jit_stats": (C.c_int, [_ip, _ip, _ip, _ip, _dp]),
    "qsv_jit_source": (C.c_int, [C.POINTER(QsvPass), C.POINTER(QsvOp), C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "qsv_jit_build_pass": (C.c_int, [C.POINTER(QsvPass), C.POINTER(QsvOp), C.c_int, C.POINTER(C.c_size_t), C.c_char_p, C.c_size_t]),
    "qsv_jit_coefs": (C.c_int, [C.POINTER(QsvPass), C.POINTER(QsvOp), C.c_int, _dp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "qsv_jit_source_scatter": (C.c_int, [C.POINTER(QsvPass), C.POINTER(QsvOp), C.c_int, C.c_int, _ip, C.c_char_p, C.c_size_t,
                                        C.POINTER(C.c_size_t)]),
    "qsv_jit_build_scatter": (C.c_int, [C.POINTER(QsvPass), C.POINTER(QsvOp), C.c_int, C.c_int, _ip, C.POINTER(C.c_size_t),
                                       C.c_char_p, C.c_size_t]),
    "qsv_norm2": (C.c_int, [_H, _dp]),
    "qsv_sample": (C.c_int, [_H, C.c_uint64, C.c_int, _dp, C.POINTER(C.c_uint64)]),
    "qsv_probabilities": (C.c_int, [_H, C.c_int, _ip, _dp]),
    "qsv_expect_z": (C.c_int, [_H, C.c_uint64, _dp]),
    "qsv_leaf_sums": (C.c_int, [_H, _dp]),
    "qsv_sample_in_leaves": (C.c_int, [_H, C.c_int, C.POINTER(C.c_uint64), _dp, _dp, C.POINTER(C.c_uint64)]),
    "qsv_comm_unique_id": (C.c_int, [C.c_void_p]),
    "qsv_comm_init": (C.c_int, [_H, C.c_void_p]),
    "qsv_swap_global_local": (C.c_int, [_H, C.c_int, _ip, _ip]),
    "qsv_comm_ipc_handle": (C.c_int, [_H, C.c_void_p]),
    "qsv_comm_set_peers": (C.c_int, [_H, C.c_void_p]),
    "qsv_comm_init_local": (C.c_int, [_H]),
    "qsv_comm_set_peers_local": (C.c_int, [_H, C.POINTER(C.c_void_p)]),
    "qsv_allreduce_sum": (C.c_int, [_H, _dp]),
    "qsv_timing_enable": (C.c_int, [_H, C.c_int]),
    "qsv_get_timings": (C.c_int, [_H, C.POINTER(QsvTiming), C.c_int, _ip]),
    "qsv_timer_start": (C.c_int, [_H]),
    "qsv_timer_stop": (C.c_int, [_H, C.POINTER(C.c_float)]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen libqsv.so and attach signatures; raises if it was not built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a). There is no CPU fallback.")
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)      # AttributeError if the .so lacks a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc: int, handle=None) -> None:
    if rc != 0:
        msg = load().qsv_last_error(handle)
        raise QsvError(rc, msg.decode(errors="replace") if msg else "?")
