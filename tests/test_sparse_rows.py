"""Amplitude-table (idx, real, imag) <-> dense state (the v1 / v2 / v3 storage format)."""
import numpy as np
import pytest

from oracle import ref_dense as O
from quantum_simulations_b200 import workloads as W
from quantum_simulations_b200.circuit.io import validate_circuit_dict
from quantum_simulations_b200.storage.sparse_rows import dense_to_rows, read_table, rows_to_dense, write_table


def test_round_trip_and_threshold():
    psi = O.simulate(validate_circuit_dict(W.ghz(6)))
    idx, re, im = dense_to_rows(psi)
    assert idx.tolist() == [0, 63] and np.allclose(re, 1 / np.sqrt(2)) and not im.any()      # GHZ: two rows
    assert np.array_equal(rows_to_dense(idx, re, im, 6), psi)
    noisy = psi + 1e-16
    assert dense_to_rows(noisy)[0].tolist() == [0, 63]                                       # 1e-15 cut of the reference
    assert np.array_equal(rows_to_dense([3, 3], [0.25, 0.25], [0, 0.5], 2), np.array([0, 0, 0, 0.5 + 0.5j]))
    with pytest.raises(ValueError):
        rows_to_dense([4], [1], [0], 2)


@pytest.mark.parametrize("suffix", [".parquet", ".csv"])
def test_files(tmp_path, suffix):
    psi = O.simulate(validate_circuit_dict(W.random_1q_cz(7, 6, 3)))
    n_rows = write_table(tmp_path / f"state{suffix}", psi)
    assert n_rows == int(((np.abs(psi.real) > 1e-15) | (np.abs(psi.imag) > 1e-15)).sum())
    got = read_table(tmp_path / f"state{suffix}", 7)
    assert np.abs(got - psi).max() <= 1e-15
