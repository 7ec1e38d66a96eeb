"""Gate fusion and level batching on the host (reference tests/test_fusion.py:21-96 restated),
plus the fused matrices the real reference produced (tests/golden/fusion_vectors.npz)."""
from pathlib import Path

import numpy as np

from oracle import ref_dense as O
from quantum_simulations_b200 import workloads as W
from quantum_simulations_b200.circuit.fusion import batch_levels, fuse_1q_ops, fusion_stats
from quantum_simulations_b200.circuit.io import levelize, validate_circuit_dict
from quantum_simulations_b200.kernel import gates as G

GOLD = Path(__file__).resolve().parent / "golden"


def test_fuse_consecutive_1q():
    H, T = G.gate_matrix("H", {}), G.gate_matrix("T", {})
    fused = fuse_1q_ops([([0], H), ([0], T)])
    assert len(fused) == 1 and fused[0][0] == [0]
    np.testing.assert_allclose(fused[0][1], T @ H, atol=1e-14)


def test_fuse_interrupted_by_2q():
    H, CX, T = G.gate_matrix("H", {}), G.gate_matrix("CNOT", {}), G.gate_matrix("T", {})
    fused = fuse_1q_ops([([0], H), ([0, 1], CX), ([0], T)])
    assert len(fused) == 3
    np.testing.assert_allclose(fused[0][1], H, atol=1e-14)
    np.testing.assert_allclose(fused[2][1], T, atol=1e-14)


def test_fuse_different_qubits_and_three_in_a_row():
    H, X, T, S = (G.gate_matrix(g, {}) for g in "HXTS")
    assert len(fuse_1q_ops([([0], H), ([1], X)])) == 2
    fused = fuse_1q_ops([([0], H), ([0], T), ([0], S)])
    assert len(fused) == 1
    np.testing.assert_allclose(fused[0][1], S @ T @ H, atol=1e-14)
    assert fuse_1q_ops([]) == []


def test_batch_all_local_and_with_nonlocal():
    levels = levelize(validate_circuit_dict(W.qft(4)))
    passes = batch_levels(levels, 4)
    assert len(passes) == 1 and passes[0]["nonlocal_ops"] == []
    assert fusion_stats(levels, 4)["fused_passes"] == 1
    passes = batch_levels(levelize(validate_circuit_dict(W.ghz(4))), 2)      # only q0, q1 local
    assert any(p["nonlocal_ops"] for p in passes)


def test_fused_step_ir_reproduces_the_state():
    """Applying the batched step IR with the oracle kernels gives the oracle's state."""
    for cd in (W.qft(6), W.ghz(6), W.random_mixed(7, 80, 2)):
        cd = validate_circuit_dict(cd)
        n = cd["number_of_qubits"]
        psi = np.zeros(1 << n, dtype=np.complex128)
        psi[0] = 1
        for step in batch_levels(levelize(cd), n):
            assert not step["nonlocal_ops"]
            O.apply_ops(psi, step["local_ops"])
        assert np.abs(psi - O.simulate(cd)).max() <= 1e-12


def test_fused_matrices_match_the_reference_golden():
    gold = np.load(GOLD / "fusion_vectors.npz")
    H, T, S, X, Y = (G.gate_matrix(g, {}) for g in "HTSXY")
    CX = G.gate_matrix("CNOT", {})
    # the op list oracle/make_golden.py fed to the reference's fuse_1q_ops
    ops = [([0], H), ([1], T), ([0], T), ([0, 1], CX), ([1], S), ([1], H), ([2], X), ([0], Y)]
    fused = fuse_1q_ops(ops)
    keys = sorted(gold.files, key=lambda k: int(k.split("/")[1]))
    assert len(keys) == len(fused)
    for key, (qs, U) in zip(keys, fused):
        assert key.split("/")[2] == "q" + "_".join(map(str, qs))
        np.testing.assert_allclose(U, gold[key], atol=1e-14)


# ------------------------------------------------------------------ fuse_2q_blocks (diagonal pair runs)
def _run(n, ops):
    psi = np.zeros(1 << n, dtype=np.complex128)
    psi[0] = 1
    for qs, U in ops:
        O.apply_1q(psi, qs[0], U) if len(qs) == 1 else O.apply_2q(psi, qs[0], qs[1], U)
    return psi


def _compiled_zz(a, b, t):
    """exp(-i t/2 Z(x)Z) the way a transpiler emits it: Rx(pi/2)-conjugated RYY ... here the common
    three-gate form with basis changes on BOTH sides, so the run opens and closes with 1-qubit gates."""
    import math
    from quantum_simulations_b200.circuit.qasm import qasm_to_ops
    txt = ('OPENQASM 2.0;\ninclude "qelib1.inc";\nqreg q[4];\n'
           f"rz(0.3) q[{a}];\nrz(-0.2) q[{b}];\ncx q[{a}],q[{b}];\nrz({t}) q[{b}];\ncx q[{a}],q[{b}];\nrz(0.11) q[{a}];\n")
    return qasm_to_ops(txt)[1]


def test_pair_runs_start_inside_the_waiting_one_qubit_gates_and_return_their_tail():
    """A Hadamard layer, then compiled ZZ blocks that share qubits: every block must come out as ONE
    diagonal gate — the run may skip the H that waits on one of its qubits only, and the 1-qubit gates
    that follow a block must be free to lead the next block on their qubit."""
    from quantum_simulations_b200.circuit.fusion import fuse_2q_blocks
    H = G.gate_matrix("H", {})
    ops = [([q], H) for q in range(4)] + _compiled_zz(0, 1, 0.7) + _compiled_zz(1, 2, -0.4) + _compiled_zz(2, 3, 1.1) \
        + _compiled_zz(0, 1, 0.2)
    fused = fuse_2q_blocks(ops, tol=1e-14)
    two = [(qs, U) for qs, U in fused if len(qs) == 2]
    assert len(two) == 4 and all(np.count_nonzero(U - np.diag(np.diag(U))) == 0 for _, U in two)
    assert len(fused) == 4 + 4                                 # the H layer + one diagonal per block
    assert np.abs(_run(4, fused) - _run(4, ops)).max() <= 1e-14


def test_pair_run_fusion_preserves_the_state_on_random_circuits():
    from quantum_simulations_b200.circuit.fusion import fuse_2q_blocks
    rng = np.random.default_rng(11)
    names1, n = ["H", "T", "S", "X", "Z"], 5
    for trial in range(60):
        ops = []
        for _ in range(int(rng.integers(10, 60))):
            k = rng.random()
            if k < 0.45:
                ops.append(([int(rng.integers(n))], G.gate_matrix(names1[int(rng.integers(len(names1)))], {})))
            elif k < 0.6:
                ops.append(([int(rng.integers(n))], G.gate_matrix("RY", {"theta": float(rng.normal())})))
            else:
                a, b = (int(x) for x in rng.choice(n, 2, replace=False))
                ops.append(([a, b], G.gate_matrix(["CNOT", "CZ", "CR"][int(rng.integers(3))], {"k": 3})))
        want = _run(n, ops)
        for kw in (dict(), dict(tol=1e-12), dict(only_diagonal=False)):
            got = _run(n, fuse_2q_blocks(ops, **kw))
            assert np.abs(got - want).max() <= 1e-12, (trial, kw)
