"""ORACLE — CPU restatement of the reference's gate-application path (NumPy, complex128).

TEST INFRASTRUCTURE ONLY.  Nothing under ``quantum_simulations_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs do, and only as the checker / the timed CPU arm.

Parity status: PINNED.  ``oracle/make_golden.py`` imports the real reference
(``/root/reference/wenbo_engine``, pure Python + NumPy, runs in the build container)
and freezes its outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks
every function here against those vectors (max|Δ| ≤ 1e-14; the 1-qubit path is
bit-identical because it evaluates the same expression per element).

What each function restates (reference file:line):
  gate_matrix        wenbo_engine/kernel/gates.py:24-108
  apply_1q           wenbo_engine/kernel/ref_dense.py:13-23  ≡ kernel/cpu_scalar.py:21-32
  apply_2q           wenbo_engine/kernel/ref_dense.py:26-41  ≡ kernel/cpu_scalar.py:35-47
  apply_1q_pair ...  wenbo_engine/kernel/cpu_nonlocal.py:22-67
  simulate           wenbo_engine/kernel/ref_dense.py:44-57
  levelize           wenbo_engine/circuit/io.py:106-117
  permute_state      wenbo_engine/circuit/staging.py:639-658

The reference materialises int64 index arrays (np.arange / fancy indexing); this
restatement addresses the same pairs/quads through strided reshape views, which is the
same arithmetic on the same elements without the 5-6x index-array memory blow-up, so it
reaches n = 28-30 on a 64 GB host.  ``apply_1q_indexed`` / ``apply_2q_indexed`` keep the
reference's gather/scatter formulation too — that one is what the CPU baseline times.
"""
from __future__ import annotations

import re

import numpy as np

C128 = np.complex128
_S2 = 1.0 / np.sqrt(2.0)


# ------------------------------------------------------------------ gate matrices
def _ctrl(u):
    m = np.eye(4, dtype=C128)
    m[2:, 2:] = u
    return m


def gate_matrix(name: str, params: dict | None = None) -> np.ndarray:
    p = params or {}
    if name == "H":
        return np.array([[_S2, _S2], [_S2, -_S2]], dtype=C128)
    if name == "X":
        return np.array([[0, 1], [1, 0]], dtype=C128)
    if name == "Y":
        return np.array([[0, -1j], [1j, 0]], dtype=C128)
    if name == "Z":
        return np.array([[1, 0], [0, -1]], dtype=C128)
    if name == "S":
        return np.array([[1, 0], [0, 1j]], dtype=C128)
    if name == "T":
        return np.array([[1, 0], [0, np.exp(1j * np.pi / 4)]], dtype=C128)
    if name == "RY":
        c, s = np.cos(p["theta"] / 2), np.sin(p["theta"] / 2)
        return np.array([[c, -s], [s, c]], dtype=C128)
    if name == "R":
        return np.array([[1, 0], [0, np.exp(2j * np.pi / 2 ** p["k"])]], dtype=C128)
    if name == "G":
        a, b = np.sqrt(1.0 / p["p"]), np.sqrt(1.0 - 1.0 / p["p"])
        return np.array([[a, -b], [b, a]], dtype=C128)
    if name == "CNOT":
        return _ctrl(gate_matrix("X"))
    if name == "CZ":
        return _ctrl(gate_matrix("Z"))
    if name == "CY":
        return _ctrl(gate_matrix("Y"))
    if name == "SWAP":
        return np.eye(4, dtype=C128)[[0, 2, 1, 3]]
    if name == "CR":
        return _ctrl(gate_matrix("R", p))
    if name == "CU":
        return _ctrl(np.linalg.matrix_power(np.asarray(p["U"], dtype=C128), p["exponent"]))
    raise ValueError(f"unknown gate {name}")


def normalise_gate(g: dict) -> tuple[str, list[int], dict]:
    """Minimal name-decoding (CR3 / R3) so the oracle accepts raw circuit dicts."""
    name = g["gate"]
    params = dict(g.get("params") or {})
    m = re.match(r"^(CR|R)(\d+)$", name)
    if m:
        name = m.group(1)
        params.setdefault("k", int(m.group(2)))
    return name, list(g["qubits"]), params


# ------------------------------------------------------------------ local kernels
def apply_1q(psi: np.ndarray, q: int, U: np.ndarray) -> None:
    """In place: for every pair (i0, i0 + 2^q) with bit q of i0 clear,
    a' = U00 a + U01 b ; b' = U10 a + U11 b."""
    if (1 << q) >= len(psi):
        raise NotImplementedError(f"qubit {q} is non-local for a chunk of {len(psi)}")
    v = psi.reshape(-1, 2, 1 << q)
    a = v[:, 0, :].copy()
    b = v[:, 1, :].copy()
    v[:, 0, :] = U[0, 0] * a + U[0, 1] * b
    v[:, 1, :] = U[1, 0] * a + U[1, 1] * b


def apply_2q(psi: np.ndarray, qa: int, qb: int, U: np.ndarray) -> None:
    """In place 4x4 on quads; sub-space row = 2*bit(qa) + bit(qb)."""
    if max(qa, qb) >= int(np.log2(len(psi))) or qa == qb:
        raise NotImplementedError("non-local or degenerate 2q gate")
    hi, lo = max(qa, qb), min(qa, qb)
    v = psi.reshape(-1, 2, 1 << (hi - lo - 1), 2, 1 << lo)

    def sel(ba: int, bb: int):
        bh, bl = (ba, bb) if qa > qb else (bb, ba)
        return v[:, bh, :, bl, :]

    order = [(0, 0), (0, 1), (1, 0), (1, 1)]          # (bit qa, bit qb) per row
    old = [sel(*o).copy() for o in order]
    for r, o in enumerate(order):
        acc = U[r, 0] * old[0]
        for c in (1, 2, 3):
            acc = acc + U[r, c] * old[c]
        sel(*o)[...] = acc


def apply_1q_indexed(psi: np.ndarray, q: int, U: np.ndarray) -> None:
    """The reference's own formulation (gather via index arrays, scatter back);
    kept so the CPU baseline pays the same memory traffic as ref_dense.py:13-23."""
    step = 1 << q
    i0 = (np.arange(0, len(psi), 2 * step)[:, None] + np.arange(step)[None, :]).ravel()
    i1 = i0 + step
    a, b = psi[i0].copy(), psi[i1].copy()
    psi[i0] = U[0, 0] * a + U[0, 1] * b
    psi[i1] = U[1, 0] * a + U[1, 1] * b


def apply_2q_indexed(psi: np.ndarray, qa: int, qb: int, U: np.ndarray) -> None:
    """Reference formulation of the quad update (ref_dense.py:26-41): boolean-select
    the bases, stack the four strided gathers, one (4x4)@(4xM) product, scatter."""
    idx = np.arange(len(psi))
    base = idx[(((idx >> qa) | (idx >> qb)) & 1) == 0]
    ia, ib = 1 << qa, 1 << qb
    sel = [base, base | ib, base | ia, base | ia | ib]
    out = U @ np.stack([psi[s] for s in sel])
    for s, row in zip(sel, out):
        psi[s] = row


# --------------------------------------------------------------- non-local kernels
def apply_1q_pair(c0, c1, U) -> None:
    a, b = c0.copy(), c1.copy()
    c0[:] = U[0, 0] * a + U[0, 1] * b
    c1[:] = U[1, 0] * a + U[1, 1] * b


def apply_2q_quad(c00, c01, c10, c11, U) -> None:
    """c01 has qb set, c10 has qa set (reference single_node.py:315-320)."""
    old = [c00.copy(), c01.copy(), c10.copy(), c11.copy()]
    for r, dst in enumerate((c00, c01, c10, c11)):
        dst[:] = sum(U[r, c] * old[c] for c in range(4))


def apply_2q_pair_qa_local(c0, c1, qa, U) -> None:
    """qa is a stride inside the chunk, qb selects the chunk (c1 = qb set)."""
    v0 = c0.reshape(-1, 2, 1 << qa)
    v1 = c1.reshape(-1, 2, 1 << qa)
    apply_2q_quad(v0[:, 0, :], v1[:, 0, :], v0[:, 1, :], v1[:, 1, :], U)


def apply_2q_pair_qb_local(c0, c1, qb, U) -> None:
    """qb is a stride inside the chunk, qa selects the chunk (c1 = qa set)."""
    v0 = c0.reshape(-1, 2, 1 << qb)
    v1 = c1.reshape(-1, 2, 1 << qb)
    apply_2q_quad(v0[:, 0, :], v0[:, 1, :], v1[:, 0, :], v1[:, 1, :], U)


# ---------------------------------------------------------------------- simulate
def simulate(circuit_dict: dict, indexed: bool = False) -> np.ndarray:
    """|0..0> through every gate in PROGRAM order (no levelize, no fusion)."""
    n = circuit_dict["number_of_qubits"]
    psi = np.zeros(1 << n, dtype=C128)
    psi[0] = 1.0
    f1, f2 = (apply_1q_indexed, apply_2q_indexed) if indexed else (apply_1q, apply_2q)
    for g in circuit_dict["gates"]:
        name, qs, params = normalise_gate(g)
        U = gate_matrix(name, params)
        if len(qs) == 1:
            f1(psi, qs[0], U)
        else:
            f2(psi, qs[0], qs[1], U)
    return psi


def apply_ops(psi: np.ndarray, ops) -> None:
    """Apply a step-IR op list [(qubits, U)] in order."""
    for qs, U in ops:
        if len(qs) == 1:
            apply_1q(psi, qs[0], U)
        else:
            apply_2q(psi, qs[0], qs[1], U)


# ----------------------------------------------------------------- host-side bits
def levelize(gates: list[dict]) -> list[list[int]]:
    """ASAP level of each gate; returns gate indices per level."""
    free: dict[int, int] = {}
    out: list[list[int]] = []
    for i, g in enumerate(gates):
        t = max((free.get(q, 0) for q in g["qubits"]), default=0)
        while len(out) <= t:
            out.append([])
        out[t].append(i)
        for q in g["qubits"]:
            free[q] = t + 1
    return out


def permute_state(state: np.ndarray, log_to_phys: list[int]) -> np.ndarray:
    """Physical layout -> logical order: out[x] = state[y], y_bit(l2p[q]) = x_bit(q)."""
    n = len(log_to_phys)
    idx = np.arange(1 << n, dtype=np.int64)
    src = np.zeros_like(idx)
    for q, p in enumerate(log_to_phys):
        src |= ((idx >> q) & 1) << p
    return state[src]


# ---------------------------------------------------- sampling (parity UNPINNED)
SAMPLE_LEAF = 1024


def sample_indices(psi: np.ndarray, seed: int, shots: int) -> np.ndarray:
    """Deterministic measurement sampling.  The reference has NO sampler
     — this definition is ours and is frozen here:
      p = re^2 + im^2 in float64; leaf sums over blocks of 1024 amplitudes accumulated
      sequentially left to right; exclusive sequential scan of leaf sums;
      u = sort(default_rng(seed).random(shots)) * total;
      result = first index i whose inclusive prefix (leaf offset + sequential in-leaf
      running sum) is > u, clamped to 2^n - 1.
    """
    p = psi.real.astype(np.float64) ** 2 + psi.imag.astype(np.float64) ** 2
    n_amp = len(p)
    leaf = min(SAMPLE_LEAF, n_amp)
    blocks = p.reshape(-1, leaf)
    leaf_sum = np.zeros(len(blocks))
    for j in range(leaf):                      # sequential accumulation inside a leaf
        leaf_sum = leaf_sum + blocks[:, j]
    offs = np.zeros(len(blocks) + 1)
    for b in range(len(blocks)):               # sequential scan over leaves
        offs[b + 1] = offs[b] + leaf_sum[b]
    total = offs[-1]
    u = np.sort(np.random.default_rng(seed).random(shots)) * total
    out = np.empty(shots, dtype=np.uint64)
    for s, x in enumerate(u):
        b = int(np.searchsorted(offs[1:], x, side="right"))
        if b >= len(blocks):
            out[s] = n_amp - 1
            continue
        run = offs[b]
        hit = leaf - 1
        for j in range(leaf):
            run = run + blocks[b, j]
            if run > x:
                hit = j
                break
        out[s] = b * leaf + hit
    return out


def marginal_probabilities(psi: np.ndarray, qubits) -> np.ndarray:
    """out[o] = sum of |psi_i|^2 over the indices i whose bit qubits[k] equals bit k of o.
    (Not in the reference: HiSVSIM-style observable; definition is ours.)"""
    p = psi.real.astype(np.float64) ** 2 + psi.imag.astype(np.float64) ** 2
    idx = np.arange(len(p), dtype=np.int64)
    o = np.zeros(len(p), dtype=np.int64)
    for k, q in enumerate(qubits):
        o |= ((idx >> q) & 1) << k
    return np.bincount(o, weights=p, minlength=1 << len(qubits))


def expectation_z(psi: np.ndarray, qubits) -> float:
    """<prod_q Z_q> = sum_i |psi_i|^2 (-1)^(number of listed qubits set in i)."""
    p = psi.real.astype(np.float64) ** 2 + psi.imag.astype(np.float64) ** 2
    idx = np.arange(len(p), dtype=np.int64)
    par = np.zeros(len(p), dtype=np.int64)
    for q in qubits:
        par ^= (idx >> q) & 1
    return float(np.sum(np.where(par == 1, -p, p)))


def run_qasm_steps(n: int, steps, cregs: dict, seed: int = 0):
    """ONE TRAJECTORY of an OpenQASM 2.0 program with mid-circuit measure / reset / classical if
    (steps of quantum_simulations_b200.circuit.qasm.qasm_to_steps).  "Parity unpinned": the reference has no
    such path (its Qiskit converter drops measurements; HiSVSIM's collapse is state_vector.hpp:829-893), so the
    definition is frozen HERE and the CUDA path (kernel.cuda_dense.run_qasm) restates it:
        rng = np.random.default_rng(seed); every measure / reset draws ONE u = rng.random(), in program order;
        p0 = sum of |amp|^2 over the indices whose bit q is 0 (float64);  outcome = 0 if u < p0 else 1;
        the other half is zeroed and the state divided by sqrt(p_outcome);  reset applies X after outcome 1;
        if(creg == value): creg read as an integer, bit 0 = least significant.
    Returns (state, {creg: int})."""
    rng = np.random.default_rng(seed)
    psi = np.zeros(1 << n, dtype=np.complex128)
    psi[0] = 1.0
    bits = {name: 0 for name in cregs}
    idx = np.arange(1 << n, dtype=np.int64)

    def measure(q: int) -> int:
        u = rng.random()
        p = marginal_probabilities(psi, [q])
        out = 0 if u < p[0] else 1
        psi[((idx >> q) & 1) != out] = 0.0
        psi[:] = psi / np.sqrt(p[out])
        return out

    for st in steps:
        if st[0] == "ops":
            apply_ops(psi, st[1])
        elif st[0] == "measure":
            _, q, creg, bit = st
            out = measure(q)
            bits[creg] = (bits[creg] & ~(1 << bit)) | (out << bit)
        elif st[0] == "reset":
            if measure(st[1]) == 1:
                apply_1q(psi, st[1], np.array([[0, 1], [1, 0]], dtype=np.complex128))
        elif st[0] == "if":
            _, creg, value, ops = st
            if bits[creg] == value:
                apply_ops(psi, ops)
        else:
            raise ValueError(f"unknown step {st[0]!r}")
    return psi, bits
