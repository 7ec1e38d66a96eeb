"""Pin the CPU oracle (oracle/) to outputs of the real reference (tests/golden/).

The golden files are produced by oracle/make_golden.py from the unmodified
wenbo_engine package; nothing here reads /root/reference.
"""
import json

import numpy as np
import pytest

from oracle import ref_dense as O
from oracle import c_oracle as C
from tests._specs import STATE_SPECS, circuit_from_spec
from tests.conftest import GOLDEN

GATE_CASES = {
    "H": {}, "X": {}, "Y": {}, "Z": {}, "S": {}, "T": {},
    "RY": {"theta": 1.2345}, "R": {"k": 3}, "G": {"p": 5},
    "CNOT": {}, "SWAP": {}, "CZ": {}, "CY": {},
    "CR": {"k": 4}, "CU": {"U": np.array([[0.6, -0.8j], [0.8j, 0.6]]), "exponent": 3},
}


@pytest.mark.parametrize("name", sorted(GATE_CASES))
def test_gate_matrices_bit_identical(golden, name):
    from quantum_simulations_b200.kernel import gates as G
    want = golden[f"gate/{name}"]
    assert np.array_equal(O.gate_matrix(name, GATE_CASES[name]), want)
    assert np.array_equal(G.gate_matrix(name, GATE_CASES[name]), want)


@pytest.mark.parametrize("spec", STATE_SPECS)
def test_simulate_matches_reference(golden, spec):
    want = golden[f"state/{spec}"]
    cd = circuit_from_spec(spec)
    for got in (O.simulate(cd), O.simulate(cd, indexed=True), C.simulate_c(cd)):
        assert got.dtype == np.complex128
        assert np.abs(got - want).max() <= 1e-14


@pytest.mark.parametrize("q", [0, 1, 4, 7])
def test_apply_1q_matches_reference_kernels(golden, q):
    U = golden["kernel/U1"]
    for fn in (O.apply_1q, O.apply_1q_indexed, C.apply_1q):
        c = golden["kernel/input"].copy()
        fn(c, q, U)
        assert np.abs(c - golden[f"kernel/scalar_1q_q{q}"]).max() <= 1e-14
        assert np.abs(c - golden[f"kernel/batched_1q_q{q}"]).max() <= 1e-14
    # same expression per element as the reference => bit identical
    c = golden["kernel/input"].copy()
    O.apply_1q(c, q, U)
    assert np.array_equal(c, golden[f"kernel/scalar_1q_q{q}"])


@pytest.mark.parametrize("qa,qb", [(0, 1), (1, 0), (2, 6), (7, 3), (7, 0)])
def test_apply_2q_matches_reference_kernels(golden, qa, qb):
    U = golden["kernel/U2"]
    for fn in (O.apply_2q, O.apply_2q_indexed, C.apply_2q):
        c = golden["kernel/input"].copy()
        fn(c, qa, qb, U)
        assert np.abs(c - golden[f"kernel/scalar_2q_{qa}_{qb}"]).max() <= 1e-13
        assert np.abs(c - golden[f"kernel/batched_2q_{qa}_{qb}"]).max() <= 1e-13


def test_nonlocal_butterflies(golden):
    chunk = golden["kernel/input"]
    parts = [chunk[i * 64:(i + 1) * 64].copy() for i in range(4)]
    U1, U2 = golden["kernel/U1"], golden["kernel/U2"]
    a, b = parts[0].copy(), parts[1].copy()
    O.apply_1q_pair(a, b, U1)
    assert np.abs(np.concatenate([a, b]) - golden["nonlocal/1q_pair"]).max() <= 1e-14
    a, b = parts[0].copy(), parts[1].copy()
    O.apply_2q_pair_qa_local(a, b, 3, U2)
    assert np.abs(np.concatenate([a, b]) - golden["nonlocal/qa_local_3"]).max() <= 1e-13
    a, b = parts[0].copy(), parts[1].copy()
    O.apply_2q_pair_qb_local(a, b, 2, U2)
    assert np.abs(np.concatenate([a, b]) - golden["nonlocal/qb_local_2"]).max() <= 1e-13
    q4 = [p.copy() for p in parts]
    O.apply_2q_quad(*q4, U2)
    assert np.abs(np.concatenate(q4) - golden["nonlocal/quad"]).max() <= 1e-13


def test_nonlocal_equals_local_on_concatenation(golden):
    """The butterfly on chunks == the local kernel on the concatenated array."""
    chunk = golden["kernel/input"]
    U2 = golden["kernel/U2"]
    whole = chunk[:128].copy()
    O.apply_2q(whole, 3, 6, U2)          # qb = 6 is the chunk bit for 64-amp chunks
    assert np.abs(whole - golden["nonlocal/qa_local_3"]).max() <= 1e-13


def test_permute_state(golden):
    got = O.permute_state(golden["permute/input"], [2, 0, 1, 4, 3])
    assert np.array_equal(got, golden["permute/l2p_2_0_1_4_3"])


def test_levelize_matches_reference():
    meta = json.loads((GOLDEN / "host_vectors.json").read_text())
    for key, want in meta.items():
        if key.startswith("levelize/"):
            cd = circuit_from_spec(key.split("/", 1)[1])
            assert O.levelize(cd["gates"]) == want


def test_non_local_qubit_raises():
    chunk = np.zeros(4, dtype=np.complex128)
    with pytest.raises(NotImplementedError, match="non-local"):
        O.apply_1q(chunk, 2, O.gate_matrix("H"))
