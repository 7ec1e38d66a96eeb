"""Circuit-dict contract — restates wenbo_engine/tests/test_contract.py:8-78 and
test_endianness_lock.py:11-12 against our host-side mirror."""
import json

import numpy as np
import pytest

from quantum_simulations_b200.circuit.io import (ENDIANNESS, validate_circuit_dict, levelize,
                                                  ALL_GATES, ALL_1Q, ALL_2Q)
from quantum_simulations_b200 import workloads as W
from tests._specs import circuit_from_spec
from tests.conftest import GOLDEN


def test_endianness_is_little():
    assert ENDIANNESS == "little"


def test_gate_sets():
    assert len(ALL_GATES) == 15 and len(ALL_1Q) == 9 and len(ALL_2Q) == 6


def test_valid_bell():
    d = validate_circuit_dict(W.bell_2q())
    assert d["number_of_qubits"] == 2
    assert [g["gate"] for g in d["gates"]] == ["H", "CNOT"]
    assert d["gates"][0]["params"] == {}


def test_valid_ry():
    d = validate_circuit_dict(W.ry_theta())
    assert d["gates"][0]["gate"] == "RY"
    assert abs(d["gates"][0]["params"]["theta"] - np.pi / 3) < 1e-12


def test_name_encoded_cr3_and_r3():
    d = validate_circuit_dict(W.cr3_encoded())
    assert d["gates"][2]["gate"] == "CR" and d["gates"][2]["params"]["k"] == 3
    d = validate_circuit_dict({"number_of_qubits": 1, "gates": [{"qubits": [0], "gate": "R7"}]})
    assert d["gates"][0] == {"qubits": [0], "gate": "R", "params": {"k": 7}}


def test_explicit_param_overrides_name_encoding():
    d = validate_circuit_dict({"number_of_qubits": 2, "gates": [
        {"qubits": [0, 1], "gate": "CR3", "params": {"k": 5}}]})
    assert d["gates"][0]["params"]["k"] == 5


@pytest.mark.parametrize("bad,stem", [
    ({"gates": []}, "missing required keys"),
    ({"number_of_qubits": 2, "gates": [{"qubits": [0], "gate": "FOOBAR"}]}, "unsupported gate"),
    ({"number_of_qubits": 2, "gates": [{"qubits": [0, 1], "gate": "H"}]}, "needs 1"),
    ({"number_of_qubits": 2, "gates": [{"qubits": [0], "gate": "CNOT"}]}, "needs 2"),
    ({"number_of_qubits": 2, "gates": [{"qubits": [5], "gate": "X"}]}, "out of range"),
    ({"number_of_qubits": 2, "gates": [], "extra": True}, "unknown top-level"),
    ({"number_of_qubits": 2, "gates": [{"qubits": [0], "gate": "RY"}]}, "requires param"),
    ({"number_of_qubits": 2, "gates": [{"qubits": [0], "gate": "H", "foo": 1}]}, "unknown keys"),
    ({"number_of_qubits": 0, "gates": []}, "positive int"),
    ({"number_of_qubits": 2, "gates": "H"}, "must be a list"),
    ({"number_of_qubits": 2, "gates": [{"qubits": [0.5], "gate": "H"}]}, r"list\[int\]"),
    ({"number_of_qubits": 2, "gates": [{"gate": "H"}]}, "missing 'qubits' or 'gate'"),
    ({"number_of_qubits": 2, "gates": [{"qubits": [0], "gate": "RY", "params": {"theta": "x"}}]},
     "bad type"),
])
def test_rejects(bad, stem):
    with pytest.raises(ValueError, match=stem):
        validate_circuit_dict(bad)


def test_not_a_dict():
    with pytest.raises(ValueError, match="must be a dict"):
        validate_circuit_dict([1, 2])


def test_levelize_and_hash_match_reference_golden():
    from quantum_simulations_b200.wal.wal import _circuit_hash
    meta = json.loads((GOLDEN / "host_vectors.json").read_text())
    for key, want in meta.items():
        kind, spec = key.split("/", 1)
        cd = validate_circuit_dict(circuit_from_spec(spec))
        if kind == "levelize":
            ids = {id(g): i for i, g in enumerate(cd["gates"])}
            assert [[ids[id(g)] for g in lv] for lv in levelize(cd)] == want
        elif kind == "circuit_hash":
            assert _circuit_hash(cd) == want


def test_levelize_counts_of_headline_workloads():
    assert len(levelize(validate_circuit_dict(W.qft(28)))) == 55
    cd = validate_circuit_dict(W.random_1q_cz(30))
    assert len(cd["gates"]) == 445 and len(levelize(cd)) == 20
    assert len(validate_circuit_dict(W.random_1q_cz(34))["gates"]) == 505
    assert len(validate_circuit_dict(W.random_1q_cz(36))["gates"]) == 535
    assert len(levelize(validate_circuit_dict(W.ghz(20)))) == 20
