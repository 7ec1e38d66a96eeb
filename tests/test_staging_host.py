"""Staging (qubit remap) on the host: restates wenbo_engine/tests/test_staging.py:53-136
(insular detection, permute_state, QubitMap) and checks every method against the oracle —
including the circuits on which the reference's own heuristic stager is wrong."""
import numpy as np
import pytest

from oracle import ref_dense as O
from quantum_simulations_b200.circuit.io import validate_circuit_dict
from quantum_simulations_b200.circuit.staging import (QubitMap, atlas_stages, non_insular_qubits,
                                                      mixed_qubits, permute_state, staging_stats)
from quantum_simulations_b200 import workloads as W


def run_steps(cd, k, method):
    steps, l2p = atlas_stages(cd, k, method=method)
    n = cd["number_of_qubits"]
    psi = np.zeros(1 << n, dtype=np.complex128)
    psi[0] = 1
    for s in steps:
        O.apply_ops(psi, s["local_ops"])      # the runner's order: local first, then non-local
        O.apply_ops(psi, s["nonlocal_ops"])
        for qs, _ in s["local_ops"]:
            assert all(q < k for q in qs)
    return permute_state(psi, l2p), steps


def test_insular_detection():
    g = lambda name, *q: {"qubits": list(q), "gate": name, "params": {}}
    for name in ("Z", "S", "T"):
        assert non_insular_qubits(g(name, 3)) == []
    assert non_insular_qubits(g("CZ", 1, 4)) == [] and non_insular_qubits(g("CR", 1, 4)) == []
    assert non_insular_qubits(g("H", 2)) == [2]
    assert non_insular_qubits(g("CNOT", 0, 3)) == [0, 3]
    assert mixed_qubits(g("CNOT", 0, 3)) == [3] and mixed_qubits(g("R", 2)) == []
    assert mixed_qubits(g("SWAP", 0, 3)) == [0, 3]


def test_permute_state_known():
    st = np.arange(4, dtype=np.complex128)
    assert np.array_equal(permute_state(st, [1, 0]), np.array([0, 2, 1, 3]))
    st = np.zeros(8, dtype=np.complex128); st[4] = 1          # physical index 4 = bit 2
    out = permute_state(st, [2, 0, 1])                          # logical q0 lives on bit 2
    assert out[1] == 1 and abs(out).sum() == 1
    assert permute_state(st, [0, 1, 2]) is st


def test_permute_state_matches_reference_golden(golden):
    got = permute_state(golden["permute/input"], [2, 0, 1, 4, 3])
    assert np.array_equal(got, golden["permute/l2p_2_0_1_4_3"])


def test_qubit_map():
    m = QubitMap(4)
    assert m.is_identity() and m.local_set(2) == {0, 1}
    m.swap_phys(0, 3)
    assert m.phys(0) == 3 and m.phys(3) == 0 and m.logical(3) == 0
    assert m.local_set(2) == {3, 1} and not m.is_identity()
    assert m.to_list() == [3, 1, 2, 0]


@pytest.mark.parametrize("method", ["heuristic", "greedy", "ilp"])
def test_reference_bug_repro(method):
    """n=3, k=1: H(2) H(0) CZ(2,0) H(0) H(1) -> the reference gives max|delta| = 0.5."""
    cd = {"number_of_qubits": 3, "gates": [
        {"qubits": [2], "gate": "H"}, {"qubits": [0], "gate": "H"}, {"qubits": [2, 0], "gate": "CZ"},
        {"qubits": [0], "gate": "H"}, {"qubits": [1], "gate": "H"}]}
    got, _ = run_steps(cd, 1, method)
    assert np.abs(got - O.simulate(validate_circuit_dict(cd))).max() <= 1e-12


@pytest.mark.parametrize("method", ["heuristic", "greedy"])
@pytest.mark.parametrize("seed", range(12))
def test_fuzz_all_methods_match_oracle(method, seed):
    n = 5 + seed % 3
    cd = W.random_mixed(n, 40, 100 + seed)
    for k in (1, 2, n - 2):
        got, _ = run_steps(cd, k, method)
        assert np.abs(got - O.simulate(validate_circuit_dict(cd))).max() <= 1e-12


@pytest.mark.parametrize("method", ["heuristic", "greedy"])
def test_named_circuits(method):
    for cd, k in ((W.ghz(5), 2), (W.qft(5), 3), (W.random_1q_cz(7, 8, 3), 4), (W.bell_2q(), 1)):
        got, _ = run_steps(cd, k, method)
        assert np.abs(got - O.simulate(validate_circuit_dict(cd))).max() <= 1e-12


def test_trivial_when_everything_is_local():
    steps, l2p = atlas_stages(W.qft(4), 4)
    assert len(steps) == 1 and l2p == [0, 1, 2, 3]


def test_stats_keys():
    s = staging_stats(W.qft(6), 3)
    assert set(s) == {"baseline_steps", "staged_steps", "baseline_nonlocal_steps",
                      "staged_nonlocal_steps", "reduction"}


def test_unknown_method():
    with pytest.raises(ValueError, match="unknown staging method"):
        atlas_stages(W.qft(4), 2, method="nope")


@pytest.mark.parametrize("seed", range(4))
def test_ilp_staging_matches_oracle(seed):
    """method="ilp" (SciPy HiGHS instead of the reference's PuLP): the plan executes to the oracle's
    state, and the ILP itself needs no more stages than the bound it was given."""
    from quantum_simulations_b200 import workloads as W
    cd = validate_circuit_dict(W.random_mixed(5, 24, seed))
    for k in (2, 3):
        got, _ = run_steps(cd, k, "ilp")
        assert np.abs(got - O.simulate(cd)).max() <= 1e-12
        assert staging_stats(cd, k, "ilp")["staged_steps"] >= 1


def test_ilp_local_sets_on_a_known_case():
    """GHZ-4 with 2 local qubits: q0,q1 first, then every later CNOT needs its own pair local; the ILP
    finds the 3-stage plan {0,1} -> {1,2} -> {2,3} (one qubit changes side per transition)."""
    from quantum_simulations_b200.circuit.staging_ilp import local_sets_ilp
    cd = validate_circuit_dict(W.ghz(4))
    sets = local_sets_ilp(cd["gates"], 4, 2, max_stages=4, time_limit=10.0)
    assert [sorted(s_) for s_ in sets] == [[0, 1], [1, 2], [2, 3]]
