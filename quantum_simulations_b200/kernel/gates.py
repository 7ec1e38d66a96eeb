"""Gate library: the 15 unitaries of the circuit contract, as complex128 ndarrays.

Mirrors the surface of the reference gate module (wenbo_engine/kernel/gates.py:24-111):
one constructor per gate, ``gate_matrix(name, params)`` as dispatcher, ``is_2q``.

Conventions (reference gates.py:3-11):
  * 1-qubit gates are 2x2.
  * 2-qubit gates are 4x4 with row/col index = 2*bit(qubits[0]) + bit(qubits[1]),
    i.e. qubits[0] is the most significant bit of the 4-dim sub-space (and is the
    control of CNOT/CY/CZ/CR/CU).

The matrices are built on the host in float64 and handed to the CUDA library as
plain ``double[8]`` / ``double[32]`` (row-major, re/im interleaved), so the GPU sees
bit-identical coefficients to the ones the CPU oracle uses.

Besides the matrices, this module classifies gates structurally
(``gate_structure``) — diagonal / controlled / permutation — which is what the pass
compiler uses to decide which qubits of a gate must sit in the on-chip tile.
"""
from __future__ import annotations

import numpy as np

_INV_SQRT2 = 1.0 / np.sqrt(2.0)
_C = np.complex128


def _m(rows) -> np.ndarray:
    return np.array(rows, dtype=_C)


# --------------------------------------------------------------------------- 1q
def H() -> np.ndarray:
    s = _INV_SQRT2
    return _m([[s, s], [s, -s]])


def X() -> np.ndarray:
    return _m([[0, 1], [1, 0]])


def Y() -> np.ndarray:
    return _m([[0, -1j], [1j, 0]])


def Z() -> np.ndarray:
    return _m([[1, 0], [0, -1]])


def S() -> np.ndarray:
    return _m([[1, 0], [0, 1j]])


def T() -> np.ndarray:
    return _m([[1, 0], [0, np.exp(1j * np.pi / 4)]])


def RY(theta: float) -> np.ndarray:
    half = theta / 2
    c, s = np.cos(half), np.sin(half)
    return _m([[c, -s], [s, c]])


def R(k: int) -> np.ndarray:
    return _m([[1, 0], [0, np.exp(2j * np.pi / 2**k)]])


def G(p: int) -> np.ndarray:
    a = np.sqrt(1.0 / p)
    b = np.sqrt(1.0 - 1.0 / p)
    return _m([[a, -b], [b, a]])


# --------------------------------------------------------------------------- 2q
def _controlled(u: np.ndarray) -> np.ndarray:
    """I (+) u in the (qubits[0], qubits[1]) big-endian sub-space."""
    out = np.eye(4, dtype=_C)
    out[2:, 2:] = u
    return out


def CNOT() -> np.ndarray:
    return _controlled(X())


def CZ() -> np.ndarray:
    return _controlled(Z())


def CY() -> np.ndarray:
    return _controlled(Y())


def SWAP() -> np.ndarray:
    out = np.zeros((4, 4), dtype=_C)
    for r, c in ((0, 0), (1, 2), (2, 1), (3, 3)):
        out[r, c] = 1
    return out


def CR(k: int) -> np.ndarray:
    return _controlled(R(k))


def CU(U, exponent: int) -> np.ndarray:
    base = np.asarray(U, dtype=_C)
    return _controlled(np.linalg.matrix_power(base, exponent))


# ------------------------------------------------------------------- dispatcher
_NO_PARAM = {"H": H, "X": X, "Y": Y, "Z": Z, "S": S, "T": T,
             "CNOT": CNOT, "SWAP": SWAP, "CZ": CZ, "CY": CY}
_WITH_PARAM = {
    "RY": lambda p: RY(p["theta"]),
    "R": lambda p: R(p["k"]),
    "G": lambda p: G(p["p"]),
    "CR": lambda p: CR(p["k"]),
    "CU": lambda p: CU(p["U"], p["exponent"]),
}
_TWO_QUBIT = frozenset({"CNOT", "SWAP", "CZ", "CY", "CR", "CU"})


def gate_matrix(name: str, params: dict) -> np.ndarray:
    """Unitary for a normalised gate entry (reference gates.py:92-108)."""
    if name in _NO_PARAM:
        return _NO_PARAM[name]()
    if name in _WITH_PARAM:
        return _WITH_PARAM[name](params)
    raise ValueError(f"unknown gate {name}")


def is_2q(name: str) -> bool:
    return name in _TWO_QUBIT


# ----------------------------------------------------------- structural classes
DIAGONAL_GATES = frozenset({"Z", "S", "T", "R", "CZ", "CR"})
CONTROLLED_GATES = frozenset({"CNOT", "CY", "CU"})  # non-diagonal only on qubits[1]


def gate_structure(name: str) -> str:
    """'diag' | 'ctrl' | 'swap' | 'dense1' — how a gate touches its qubits.

    diag  : diagonal in the computational basis; no qubit needs to be in the tile
            (a per-amplitude phase that depends only on index bits).
    ctrl  : |0><0| (x) I + |1><1| (x) u ; only qubits[1] (the target) is mixed.
    swap  : a relabelling of two index bits.
    dense1: a general 1-qubit gate; its qubit is mixed.
    """
    if name in DIAGONAL_GATES:
        return "diag"
    if name in CONTROLLED_GATES:
        return "ctrl"
    if name == "SWAP":
        return "swap"
    return "dense1"
