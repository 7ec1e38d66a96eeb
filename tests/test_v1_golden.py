"""The reference's FIRST implementation (v1: SQLite amplitude tables, SQL JOIN + GROUP BY per gate) as a second,
independent pin: tests/golden/v1_sqlite_states.json holds the rows (idx, real, imag) it produced for its own
benchmark circuits (oracle/make_golden_v1.py ran the unmodified v1_implementation/src).  The oracle restatement,
the row importer / exporter and the CUDA path must all reproduce them."""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import ref_dense as O
from quantum_simulations_b200.circuit.io import validate_circuit_dict
from quantum_simulations_b200.storage import sparse_rows as SR

GOLD = json.loads((Path(__file__).resolve().parent / "golden" / "v1_sqlite_states.json").read_text())


def _circuit(case):
    cd = {"number_of_qubits": case["circuit"]["number_of_qubits"], "gates": []}
    for g in case["circuit"]["gates"]:
        p = dict(g["params"])
        if "U" in p:
            p["U"] = np.array(p["U"]["re"]) + 1j * np.array(p["U"]["im"])
        p.pop("name", None)
        cd["gates"].append({"qubits": g["qubits"], "gate": g["gate"], "params": p})
    return cd


def _dense(case):
    n = case["circuit"]["number_of_qubits"]
    psi = np.zeros(1 << n, dtype=np.complex128)
    for i, re, im in case["rows"]:
        psi[i] = re + 1j * im
    return psi


@pytest.mark.parametrize("name", sorted(GOLD))
def test_oracle_reproduces_the_sqlite_implementation(name):
    want = _dense(GOLD[name])
    got = O.simulate(validate_circuit_dict(_circuit(GOLD[name])))
    assert abs(np.vdot(want, want).real - 1.0) < 1e-12
    assert np.abs(got - want).max() <= 1e-12


@pytest.mark.parametrize("name", sorted(GOLD))
def test_row_format_round_trip(name):
    """storage/sparse_rows: (idx, real, imag) rows <-> dense, the table format of v1 / v2 / v3."""
    case = GOLD[name]
    n = case["circuit"]["number_of_qubits"]
    idx = np.array([r[0] for r in case["rows"]], dtype=np.int64)
    re = np.array([r[1] for r in case["rows"]]); im = np.array([r[2] for r in case["rows"]])
    dense = SR.rows_to_dense(idx, re, im, n)
    assert np.array_equal(dense, _dense(case))
    i2, r2, m2 = SR.dense_to_rows(dense, tol=0.0)
    keep = (re != 0) | (im != 0)
    assert np.array_equal(i2, idx[keep]) and np.array_equal(r2, re[keep]) and np.array_equal(m2, im[keep])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLD))
def test_cuda_path_reproduces_the_sqlite_implementation(name):
    from quantum_simulations_b200.kernel.cuda_dense import simulate
    cd = validate_circuit_dict(_circuit(GOLD[name]))
    want = _dense(GOLD[name])
    for fused in (True, False):
        assert np.abs(simulate(cd, fused=fused) - want).max() <= 1e-12
    assert np.abs(simulate(cd, dtype="complex64") - want).max() <= 1e-5
