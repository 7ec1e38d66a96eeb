"""ctypes front for oracle/ref_dense_c.c (TEST INFRASTRUCTURE ONLY).

`simulate_c` drives the C pair/quad loops gate by gate in program order, exactly like
oracle.ref_dense.simulate, but multi-threaded and without index arrays — it is the
"all host cores" CPU arm of bench.py and the large-n checker.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

from oracle.ref_dense import gate_matrix, normalise_gate

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "liboracle_c.so"
_lib = None


def build(force: bool = False) -> Path:
    if force or not _SO.exists() or _SO.stat().st_mtime < (_HERE / "ref_dense_c.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-s", "-B"], check=True)
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not _SO.exists():
            build()
        L = ctypes.CDLL(str(_SO))
        dp = ctypes.POINTER(ctypes.c_double)
        L.oracle_apply_1q.argtypes = [dp, ctypes.c_int, ctypes.c_int, dp]
        L.oracle_apply_2q.argtypes = [dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, dp]
        L.oracle_init_zero.argtypes = [dp, ctypes.c_int]
        L.oracle_norm2.argtypes = [dp, ctypes.c_int]
        L.oracle_norm2.restype = ctypes.c_double
        _lib = L
    return _lib


def _dp(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def n_threads() -> int:
    return int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))


def apply_1q(psi: np.ndarray, q: int, U: np.ndarray) -> None:
    n = int(np.log2(len(psi)))
    u = np.ascontiguousarray(U, dtype=np.complex128)
    lib().oracle_apply_1q(_dp(psi), n, q, _dp(u))


def apply_2q(psi: np.ndarray, qa: int, qb: int, U: np.ndarray) -> None:
    n = int(np.log2(len(psi)))
    u = np.ascontiguousarray(U, dtype=np.complex128)
    lib().oracle_apply_2q(_dp(psi), n, qa, qb, _dp(u))


def simulate_c(circuit_dict: dict, psi: np.ndarray | None = None) -> np.ndarray:
    n = circuit_dict["number_of_qubits"]
    if psi is None:
        psi = np.empty(1 << n, dtype=np.complex128)
        lib().oracle_init_zero(_dp(psi), n)
    for g in circuit_dict["gates"]:
        name, qs, params = normalise_gate(g)
        U = gate_matrix(name, params)
        if len(qs) == 1:
            apply_1q(psi, qs[0], U)
        else:
            apply_2q(psi, qs[0], qs[1], U)
    return psi
