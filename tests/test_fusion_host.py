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
