"""The planner keeps every 2^11-amplitude pass inside what a run-time specialised kernel holds (PassCompiler(fit_jit=True):
csrc/jit.cuh kMaxOps = 256 micro-ops, kMaxCoefs = 448 coefficients), so that deep circuits do not fall back to the
interpreting kernels.  CPU only: the library's generator is asked whether it accepts each pass (qsv_jit_source /
qsv_jit_coefs), and the re-planned programs run through the NumPy pass emulator against the oracle."""
import ctypes as C

import numpy as np
import pytest

from oracle import ref_dense as O
from quantum_simulations_b200 import _lib as L
from quantum_simulations_b200 import workloads as W
from quantum_simulations_b200.circuit import sharding
from quantum_simulations_b200.circuit.io import validate_circuit_dict
from quantum_simulations_b200.circuit.passes import (JIT_MAX_COEFS, JIT_MAX_OPS, PassCompiler, fits_jit, jit_coefs)
from quantum_simulations_b200.kernel import gates as G
from tests.pass_emulator import run_program, run_program_sharded


def _ops(cd):
    return [(q["qubits"], G.gate_matrix(q["gate"], q["params"])) for q in cd["gates"]]


def _deep_rotations(n: int, layers: int, seed: int) -> dict:
    """RY on every qubit + a CNOT ladder per layer: rotation-heavy like QASMBench's dnn_n16 (4 coefficients per
    absorbed rotation), the shape that overflowed the coefficient bank."""
    rng = np.random.default_rng(seed)
    gates = []
    for _ in range(layers):
        for q in range(n):
            gates.append({"gate": "RY", "qubits": [q], "params": {"theta": float(rng.uniform(0, 2 * np.pi))}})
            gates.append({"gate": "R", "qubits": [q], "params": {"k": int(rng.integers(1, 6))}})
        for q in range(0, n - 1):
            if q % 4 != 3:                      # ladders inside groups of four qubits: a round of a pass swallows many layers
                gates.append({"gate": "CNOT", "qubits": [q, q + 1], "params": {}})
    return validate_circuit_dict({"number_of_qubits": n, "gates": gates})


def _library_accepts(step, dtype=L.QSV_C128):
    need = C.c_size_t()
    out = (C.c_double * 2048)()
    rc = L.load().qsv_jit_coefs(C.byref(step.desc), step.ops, dtype, out, 2048, C.byref(need))
    return rc == 0, need.value


@pytest.mark.parametrize("workload", ["random", "qft", "mixed", "deep"])
def test_coefficient_count_is_the_generators(workload):
    n = 13
    cd = {"random": lambda: W.random_1q_cz(n, 12, 3), "qft": lambda: W.qft(n), "mixed": lambda: W.random_mixed(n, 300, 5),
          "deep": lambda: _deep_rotations(n, 6, 1)}[workload]()
    prog = PassCompiler(n, n, fit_jit=False, max_ops=200).compile(_ops(validate_circuit_dict(cd)))
    seen = 0
    for s in prog.passes:
        ok, need = _library_accepts(s)
        if ok:                                  # (a refused pass reports no count)
            assert need == jit_coefs(s)
            seen += 1
    assert seen


def test_deep_circuit_without_the_budget_overflows_and_with_it_fits():
    n = 13
    cd = _deep_rotations(n, 14, 2)
    loose = PassCompiler(n, n, fit_jit=False).compile(_ops(cd))
    assert any(not fits_jit(s) for s in loose.passes), "the case must actually exceed the limits to test anything"
    assert any(not _library_accepts(s)[0] for s in loose.passes)
    prog = PassCompiler(n, n).compile(_ops(cd))
    assert prog.stats["passes_beyond_jit"] == 0 and prog.stats["fit_jit_max_ops"] < 380
    for s in prog.passes:
        assert s.desc.n_ops <= JIT_MAX_OPS and jit_coefs(s) <= JIT_MAX_COEFS
        assert _library_accepts(s)[0]
    psi = np.zeros(1 << n, dtype=np.complex128)
    psi[0] = 1
    assert np.abs(run_program(prog, psi) - O.simulate(cd)).max() <= 1e-11


def test_planned_and_sharded_deep_circuits_fit_and_match_the_oracle():
    n = 13
    cd = _deep_rotations(n, 10, 4)
    single = sharding.plan_single(_ops(cd), n, "complex128", True, False)
    assert all(fits_jit(s) for s in single.passes)
    psi = np.random.default_rng(0).standard_normal(1 << n) + 0j if single.fused_init else None
    if psi is None:
        psi = np.zeros(1 << n, dtype=np.complex128)
        psi[0] = 1
    assert np.abs(run_program(single, psi) - O.simulate(cd)).max() <= 1e-11
    sh = sharding.plan(_ops(cd), n, n - 1, swap_anywhere=True, rank_flips=True)
    assert all(fits_jit(s) for s in sh.passes)
    psi = np.zeros(1 << n, dtype=np.complex128)
    psi[0] = 1
    if sh.fused_init:
        psi = np.random.default_rng(1).standard_normal(1 << n) + 0j
    out = run_program_sharded(sh, psi)
    if sh.rank_flip_mask:
        shards = out.reshape(2, -1)
        out = np.concatenate([shards[r ^ sh.rank_flip_mask] for r in range(2)])
    assert np.abs(out - O.simulate(cd)).max() <= 1e-11


def test_baseline_workloads_are_planned_once_as_before():
    """The BASELINE circuits fit: fit_jit must not change a single pass of their plans."""
    for n, cd in ((20, W.random_1q_cz(20, 20, 1234)), (18, W.qft(18)), (16, W.ghz(16))):
        ops = _ops(validate_circuit_dict(cd))
        a = sharding.plan_single(ops, n, "complex128", True, False, max_rounds=3)
        b = sharding.plan_single(ops, n, "complex128", True, False, max_rounds=3, fit_jit=False)
        assert "fit_jit_max_ops" not in a.stats
        assert [bytes(s.desc) for s in a.passes] == [bytes(s.desc) for s in b.passes]
