"""HiSVSIM part-file import (circuit/hisvsim_parts.py)."""
import numpy as np
import pytest

from oracle import ref_dense as O
from quantum_simulations_b200.circuit.hisvsim_parts import qasm_with_parts, read_part_file, reorder_by_parts
from quantum_simulations_b200.circuit.qasm import qasm_to_ops

QASM = """OPENQASM 2.0;
include "qelib1.inc";
qreg q[4];
creg c[4];
h q[0];
h q[2];
cx q[0],q[1];
cx q[2],q[3];
rz(0.3) q[1];
ccx q[0],q[1],q[2];
h q[3];
measure q -> c;
"""
# the format of the vendored files: entry nodes (index 0), gates (1-based), exit nodes
PARTS = """0 q0 0
0 q1 0
0 q2 1
0 q3 1
1 h_0 0
2 h_2 1
3 cx_4 0
4 cx_7 1
5 rz_10 0
6 ccx_12 2
7 h_16 2
8293 q0_exit_18 2
8293 q3_exit_19 2
"""


def state(n, ops):
    psi = np.zeros(1 << n, dtype=np.complex128)
    psi[0] = 1
    O.apply_ops(psi, ops)
    return psi


def test_read_and_reorder():
    assert read_part_file(PARTS) == [0, 1, 0, 1, 0, 2, 2]
    n, ops = qasm_to_ops(QASM)
    n2, new_ops, qsets = qasm_with_parts(QASM, PARTS)
    assert n2 == n == 4 and len(new_ops) == len(ops)
    assert qsets == [{0, 1}, {2, 3}, {0, 1, 2, 3}]
    # part 0's gates first, in circuit order
    assert [q for q, _ in new_ops[:3]] == [[0], [0, 1], [1]]
    assert np.abs(state(n, new_ops) - state(n, ops)).max() < 1e-15


def test_cyclic_partition_is_rejected():
    _, ops = qasm_to_ops('OPENQASM 2.0; qreg q[2]; h q[0]; cx q[0],q[1]; h q[0];')
    with pytest.raises(ValueError, match="cyclic"):
        reorder_by_parts(ops, [0, 1, 0])
    with pytest.raises(ValueError, match="part labels"):
        reorder_by_parts(ops, [0, 1])
    with pytest.raises(ValueError, match="1..G"):
        read_part_file("1 h_0 0\n3 h_2 0\n")
