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


def test_parts_run_as_stages():
    """sharding.plan_parts: every part of a partition is ONE stage of the sharded program — a single gather of the
    qubits it mixes in front of it, then passes without communication (HiSVSIM execute.hpp:665-685).  Checked
    through the pass emulator on 2/4/8 shards against the oracle, for the part file above and for partitions
    cut by the engine's own 'nat' splitter."""
    from quantum_simulations_b200 import workloads as W
    from quantum_simulations_b200.circuit import sharding
    from quantum_simulations_b200.circuit.hisvsim_parts import qasm_parts
    from quantum_simulations_b200.circuit.io import validate_circuit_dict
    from quantum_simulations_b200.circuit.passes import PassStep, SwapStep
    from quantum_simulations_b200.kernel import gates as G
    from tests.pass_emulator import run_program_sharded

    def run(prog, n):
        psi = np.random.default_rng(1).standard_normal(1 << n) + 0j          # fused init ignores what is there
        if not prog.fused_init:
            psi = np.zeros(1 << n, dtype=np.complex128)
            psi[0] = 1
        return run_program_sharded(prog, psi)

    # the QASMBench-style file with its part file, padded to 9 qubits so that it can be sharded 2 ways
    qasm9 = QASM.replace("qreg q[4];", "qreg q[4]; qreg pad[5];").replace("creg c[4];", "creg c[4]; creg d[5];") \
                .replace("measure q -> c;", "h pad[4]; cx pad[4],q[0]; measure q -> c;")
    parts9 = PARTS.replace("7 h_16 2\n", "7 h_16 2\n8 h_17 3\n9 cx_18 3\n")
    n, parts = qasm_parts(qasm9, parts9)
    assert n == 9 and [len(p) for p in parts] == [3, 2, 16, 2]               # ccx expands to 15 ops
    prog = sharding.plan_parts(parts, n, n - 1, tile_bits=6, low_bits=2, swap_anywhere=True)
    want = state(n, [op for p in parts for op in p])
    assert np.abs(run(prog, n) - want).max() <= 1e-12
    for n, g, cd in ((10, 2, W.random_mixed(10, 150, 3)), (11, 1, W.qft(11)), (12, 3, W.random_1q_cz(12, 12, 7))):
        cd = validate_circuit_dict(cd)
        ops = [(q["qubits"], G.gate_matrix(q["gate"], q["params"])) for q in cd["gates"]]
        parts = sharding.split_into_parts(ops, n - g - 1)
        assert all(len(set().union(*[sharding.mixed_qubits(qs, U) for qs, U in p])) <= n - g - 1 for p in parts)
        prog = sharding.plan_parts(parts, n, n - g, tile_bits=6, low_bits=2, swap_anywhere=True)
        assert prog.final_pos == list(range(n)) and prog.stats["parts"] == len(parts)
        # at most one swap step in front of every part and one to restore the identity layout at the end
        assert sum(isinstance(s_, SwapStep) for s_ in prog.steps) <= len(parts) + 1
        assert np.abs(run(prog, n) - O.simulate(cd)).max() <= 1e-12
    with pytest.raises(ValueError, match="cannot run as one stage"):
        sharding.plan_parts([ops], 12, 9)
