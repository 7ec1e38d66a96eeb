"""Workload definitions (circuit dicts) used by tests, smoke and bench.

* ``bell_2q, x_on_q0_3q, ry_theta, cr3_encoded, ghz, qft`` restate the reference
  fixtures (wenbo_engine/tests/fixtures/circuits.py:7-63; GHZ/QFT are the same circuits
  as v1_implementation/src/circuits.py) so our parity tests read like the reference's.
* ``random_1q_cz`` is the seeded "random depth-d (1q + CZ layers)" family that
  BASELINE.json's configs 3-5 name.  The reference has no such generator; the definition
  is frozen in SURVEY.md §8(d) and restated here.
"""
from __future__ import annotations

import numpy as np


def _g(name: str, *qubits: int, **params) -> dict:
    entry = {"qubits": list(qubits), "gate": name}
    if params:
        entry["params"] = params
    return entry


def bell_2q() -> dict:
    return {"number_of_qubits": 2, "gates": [_g("H", 0), _g("CNOT", 0, 1)]}


def x_on_q0_3q() -> dict:
    return {"number_of_qubits": 3, "gates": [_g("X", 0)]}


def ry_theta() -> dict:
    return {"number_of_qubits": 2, "gates": [_g("RY", 0, theta=np.pi / 3)]}


def cr3_encoded() -> dict:
    return {"number_of_qubits": 2, "gates": [_g("H", 0), _g("H", 1), _g("CR3", 0, 1)]}


def ghz(n: int) -> dict:
    """H(0) then a CNOT chain (q-1 -> q): (|0..0> + |1..1>)/sqrt2."""
    return {"number_of_qubits": n,
            "gates": [_g("H", 0)] + [_g("CNOT", q - 1, q) for q in range(1, n)]}


def qft(n: int) -> dict:
    """Textbook QFT without the final swaps: H(j); CR_{k-j+1}(control k, target j), k>j."""
    gates = []
    for j in range(n):
        gates.append(_g("H", j))
        gates.extend(_g("CR", k, j, k=k - j + 1) for k in range(j + 1, n))
    return {"number_of_qubits": n, "gates": gates}


def hadamard_wall(n: int) -> dict:
    return {"number_of_qubits": n, "gates": [_g("H", q) for q in range(n)]}


_RANDOM_1Q = ("H", "X", "Y", "S", "T", "RY")


def random_1q_cz(n: int, depth: int = 20, seed: int = 1234) -> dict:
    """Seeded random circuit of alternating 1-qubit and brickwork-CZ layers.

    Layer d even: for q = 0..n-1 one gate drawn uniformly from {H,X,Y,S,T,RY(theta)}
    (draw order per qubit: the gate index ``rng.integers(6)``, then
    ``theta = rng.uniform(0, 2*pi)`` only if RY).
    Layer d odd : CZ on (q, q+1) for all q with q % 2 == ((d-1)//2) % 2 and q+1 < n.
    n=30, depth=20 -> 300 one-qubit gates + 145 CZ = 445 gates in 20 levels.
    """
    rng = np.random.default_rng(seed)
    gates: list[dict] = []
    for d in range(depth):
        if d % 2 == 0:
            for q in range(n):
                name = _RANDOM_1Q[int(rng.integers(6))]
                if name == "RY":
                    gates.append(_g("RY", q, theta=float(rng.uniform(0.0, 2.0 * np.pi))))
                else:
                    gates.append(_g(name, q))
        else:
            start = ((d - 1) // 2) % 2
            gates.extend(_g("CZ", q, q + 1) for q in range(start, n - 1, 2))
    return {"number_of_qubits": n, "gates": gates}


def random_mixed(n: int, n_gates: int, seed: int) -> dict:
    """Seeded circuit over ALL 15 gate names with random qubits — the parity fuzzer."""
    rng = np.random.default_rng(seed)
    one = ("H", "X", "Y", "Z", "S", "T", "RY", "R", "G")
    two = ("CNOT", "SWAP", "CZ", "CY", "CR", "CU")
    gates: list[dict] = []
    for _ in range(n_gates):
        if n >= 2 and rng.random() < 0.45:
            name = two[int(rng.integers(len(two)))]
            a, b = (int(x) for x in rng.choice(n, size=2, replace=False))
            if name == "CR":
                gates.append(_g("CR", a, b, k=int(rng.integers(1, 6))))
            elif name == "CU":
                th, ph = rng.uniform(0, 2 * np.pi, size=2)
                u = [[np.cos(th), -np.exp(1j * ph) * np.sin(th)],
                     [np.exp(-1j * ph) * np.sin(th), np.cos(th)]]
                gates.append(_g("CU", a, b, U=u, exponent=int(rng.integers(1, 4))))
            else:
                gates.append(_g(name, a, b))
        else:
            name = one[int(rng.integers(len(one)))]
            q = int(rng.integers(n))
            if name == "RY":
                gates.append(_g("RY", q, theta=float(rng.uniform(0, 2 * np.pi))))
            elif name == "R":
                gates.append(_g("R", q, k=int(rng.integers(1, 6))))
            elif name == "G":
                gates.append(_g("G", q, p=int(rng.integers(2, 9))))
            else:
                gates.append(_g(name, q))
    return {"number_of_qubits": n, "gates": gates}
