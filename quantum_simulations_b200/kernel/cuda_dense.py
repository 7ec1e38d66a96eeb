"""GPU counterpart of the reference's in-memory simulator
(wenbo_engine/kernel/ref_dense.py:44-57): ``simulate(circuit_dict) -> ndarray``.

Same contract — validate, start from |0...0>, apply every gate, return the final state as a
host ndarray in logical qubit order — but the gates run as fused on-chip passes on the B200
(circuit/passes.py + csrc/pass_kernel.cuh).  This is the call bench.py times end to end.
"""
from __future__ import annotations

import numpy as np

from quantum_simulations_b200.circuit.io import validate_circuit_dict
from quantum_simulations_b200.circuit.passes import PassCompiler, Program, REG_BITS
from quantum_simulations_b200.kernel import gates as gmod


def circuit_ops(cd: dict) -> list:
    """Normalised circuit -> step-IR op list [(qubits, U)] in program order."""
    return [(g["qubits"], gmod.gate_matrix(g["gate"], g["params"])) for g in cd["gates"]]


def compile_circuit(circuit_dict: dict, dtype="complex128", zero_init: bool = True, skip_zero_support: bool = False,
                    **compiler_kw) -> Program:
    """Passes for a run of the whole circuit.  zero_init=True (the run starts from |0...0>, like the
    reference's simulate): the initial qubit placement is free and is chosen so that no
    layout-restoring pass is needed (circuit/sharding.plan_single)."""
    from quantum_simulations_b200.circuit.sharding import plan_single
    cd = validate_circuit_dict(circuit_dict)
    return plan_single(circuit_ops(cd), cd["number_of_qubits"], np.dtype(dtype).name, zero_init,
                       skip_zero_support, **compiler_kw)


def simulate(circuit_dict: dict, dtype="complex128", device: int = 0, out: np.ndarray | None = None,
             fused: bool = True, jit: bool | None = None, phases: dict | None = None,
             skip_zero_support: bool = False, **compiler_kw) -> np.ndarray:
    """Run the circuit on the GPU and return the final state vector (host, `dtype`).
    `phases` (optional dict) receives the host time of each phase in milliseconds.
    skip_zero_support=True: the run starts from |0...0>, so index bits of qubits no pass has touched
    yet are 0 everywhere; passes visit only the tiles that can hold data (exact: zero tiles stay
    zero).  Off by default so that every pass streams the whole state (the roofline accounting)."""
    import time
    from quantum_simulations_b200.kernel.cuda import DeviceState

    t = [time.perf_counter()]

    def mark(name):
        if phases is not None:
            t.append(time.perf_counter())
            phases[name] = phases.get(name, 0.0) + (t[-1] - t[-2]) * 1e3

    cd = validate_circuit_dict(circuit_dict)
    n = cd["number_of_qubits"]
    ops = circuit_ops(cd)
    mark("validate+gate_matrices")
    with DeviceState(n, dtype, device) as st:
        mark("create(cudaMalloc)")
        if fused and n >= REG_BITS:
            from quantum_simulations_b200.circuit.sharding import plan_single
            prog = plan_single(ops, n, st.dtype.name, True, skip_zero_support, **compiler_kw)
            mark("pass_compiler")
            if not prog.fused_init:            # otherwise the first pass creates |0...0> itself
                st.init_zero()
            st.run_program(prog, jit=jit)
            if phases is not None:
                st.sync()
            mark("upload+specialise+run")
        else:
            st.init_zero()
            for qs, U in ops:
                st.apply_op(qs, U)
            if phases is not None:
                st.sync()
            mark("per_gate_run")
        res = st.download(out)
        mark("download(D2H)")
    mark("destroy(cudaFree)")
    return res


def run_qasm(text: str, seed: int = 0, dtype="complex128", device: int = 0, **compiler_kw):
    """ONE TRAJECTORY of an OpenQASM 2.0 program with mid-circuit `measure`, `reset` and `if(c==k)` on the GPU:
    unitary segments run as fused programs on the resident state, every measurement is a marginal
    (qsv_probabilities) + a seeded draw + a projection sweep (DeviceState.project; HiSVSIM
    state_vector.hpp:829-893).  Semantics, draw order and the outcome rule are those of
    oracle/ref_dense.py::run_qasm_steps.  Returns (final state, {creg: int})."""
    from quantum_simulations_b200.circuit.fusion import fuse_2q_blocks
    from quantum_simulations_b200.circuit.qasm import qasm_to_steps
    from quantum_simulations_b200.circuit.sharding import plan_single
    from quantum_simulations_b200.kernel.cuda import DeviceState

    n, steps, cregs = qasm_to_steps(text)
    rng = np.random.default_rng(seed)
    bits = {name: 0 for name in cregs}
    X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
    with DeviceState(n, dtype, device) as st:
        st.init_zero()

        def run_ops(ops) -> None:
            ops = fuse_2q_blocks(ops, tol=1e-14)
            if not ops:
                return
            if n >= REG_BITS:
                st.run_program(plan_single(ops, n, st.dtype.name, False, False, **compiler_kw))   # from the state as it is
            else:
                for qs, U in ops:
                    st.apply_op(qs, U)

        def measure(q: int) -> int:
            u = rng.random()
            p = st.probabilities([q])
            outcome = 0 if u < p[0] else 1
            st.project(q, outcome)
            return outcome

        for step in steps:
            if step[0] == "ops":
                run_ops(step[1])
            elif step[0] == "measure":
                _, q, creg, bit = step
                bits[creg] = (bits[creg] & ~(1 << bit)) | (measure(q) << bit)
            elif step[0] == "reset":
                if measure(step[1]) == 1:
                    st.apply_1q(step[1], X)
            elif step[0] == "if":
                if bits[step[1]] == step[2]:
                    run_ops(step[3])
            else:
                raise ValueError(f"unknown step {step[0]!r}")
        return st.download(), bits


def simulate_qasm(text: str, dtype="complex128", device: int = 0, out: np.ndarray | None = None, **compiler_kw) -> np.ndarray:
    """OpenQASM 2.0 program -> final state on the GPU (circuit/qasm.py front end; CNOT-ladder
    decompositions of controlled phases are fused back into diagonal blocks first).  Programs with
    mid-circuit measure / reset / if are not unitary: use run_qasm (one seeded trajectory)."""
    from quantum_simulations_b200.circuit.fusion import fuse_2q_blocks
    from quantum_simulations_b200.circuit.qasm import qasm_to_ops
    from quantum_simulations_b200.kernel.cuda import DeviceState

    n, ops = qasm_to_ops(text)
    ops = fuse_2q_blocks(ops, tol=1e-14)
    with DeviceState(n, dtype, device) as st:
        if n >= REG_BITS:
            from quantum_simulations_b200.circuit.sharding import plan_single
            prog = plan_single(ops, n, st.dtype.name, True, False, **compiler_kw)    # as simulate(): free initial
            if not prog.fused_init:                                                  # placement, fused |0...0>
                st.init_zero()
            st.run_program(prog)
        else:
            st.init_zero()
            for qs, U in ops:
                st.apply_op(qs, U)
        return st.download(out)
