"""OpenQASM 2.0 front end (circuit/qasm.py): parsed programs against the oracle."""
import math

import numpy as np
import pytest

from oracle import ref_dense as O
from quantum_simulations_b200 import workloads as W
from quantum_simulations_b200.circuit.io import validate_circuit_dict
from quantum_simulations_b200.circuit.passes import PassCompiler
from quantum_simulations_b200.circuit.qasm import QasmError, qasm_to_dict, qasm_to_ops
from tests.pass_emulator import run_program

HEAD = 'OPENQASM 2.0;\ninclude "qelib1.inc";\n'


def state(n, ops):
    psi = np.zeros(1 << n, dtype=np.complex128)
    psi[0] = 1
    O.apply_ops(psi, ops)
    return psi


def qft_qasm(n):
    """The decomposition QASMBench uses (qft_n20.qasm: u1 / cx ladders), on qubits in our order."""
    out = [HEAD, f"qreg q[{n}];", f"creg c[{n}];"]
    for j in range(n):
        out.append(f"h q[{j}];")
        for k in range(j + 1, n):
            lam = 2 * math.pi / 2 ** (k - j + 1)
            out += [f"u1({lam / 2}) q[{k}];", f"cx q[{k}],q[{j}];", f"u1({-lam / 2}) q[{j}];", f"cx q[{k}],q[{j}];",
                    f"u1({lam / 2}) q[{j}];"]
    out.append("measure q -> c;")
    return "\n".join(out)


def test_bell_and_dict_form():
    src = HEAD + "qreg q[2]; creg c[2];\nh q[0];\ncx q[0],q[1]; // entangle\nbarrier q;\nmeasure q[0] -> c[0];"
    n, ops = qasm_to_ops(src)
    assert n == 2 and np.abs(state(n, ops) - np.array([1, 0, 0, 1]) / np.sqrt(2)).max() < 1e-15
    cd = qasm_to_dict(src)
    assert cd == {"number_of_qubits": 2, "gates": [{"qubits": [0], "gate": "H", "params": {}},
                                                   {"qubits": [0, 1], "gate": "CNOT", "params": {}}]}
    assert np.abs(O.simulate(validate_circuit_dict(cd)) - state(n, ops)).max() < 1e-15


def test_qasmbench_style_qft_equals_the_reference_qft():
    n = 6
    _, ops = qasm_to_ops(qft_qasm(n))
    want = O.simulate(validate_circuit_dict(W.qft(n)))
    assert np.abs(state(n, ops) - want).max() < 1e-12
    # and through the pass compiler (emulated passes)
    psi = np.zeros(1 << n, dtype=np.complex128)
    psi[0] = 1
    run_program(PassCompiler(n, tile_bits=5, low_bits=2).compile(ops), psi)
    assert np.abs(psi - want).max() < 1e-12


def test_two_qubit_fusion_recovers_controlled_phases():
    """cx / u1 ladders fuse back into diagonal 2-qubit blocks: same state, far fewer mixing ops."""
    from quantum_simulations_b200.circuit.fusion import fuse_2q_blocks
    n = 8
    _, ops = qasm_to_ops(qft_qasm(n))
    fused = fuse_2q_blocks(ops)
    want = O.simulate(validate_circuit_dict(W.qft(n)))
    assert np.abs(state(n, fused) - want).max() < 1e-12
    assert len(fused) < len(ops) / 2
    two_q = [U for q, U in fused if len(q) == 2]
    assert two_q and all(not np.any(U - np.diag(np.diag(U))) for U in two_q)                   # diagonal again
    plain = PassCompiler(n, tile_bits=6, low_bits=2).compile(ops)
    better = PassCompiler(n, tile_bits=6, low_bits=2).compile(fused)
    psi = np.zeros(1 << n, dtype=np.complex128)
    psi[0] = 1
    run_program(better, psi)
    assert np.abs(psi - want).max() < 1e-12
    assert better.stats["micro_ops"] < plain.stats["micro_ops"] / 2
    dense = fuse_2q_blocks(ops, only_diagonal=False)                 # every pair run -> one 4x4 block
    assert np.abs(state(n, dense) - want).max() < 1e-12 and len(dense) < len(fused)
    # random mixed circuits: fusion never changes the state
    for seed in range(5):
        cd = validate_circuit_dict(W.random_mixed(7, 120, seed))
        from quantum_simulations_b200.kernel import gates as G
        ir = [(g["qubits"], G.gate_matrix(g["gate"], g["params"])) for g in cd["gates"]]
        assert np.abs(state(7, fuse_2q_blocks(ir)) - O.simulate(cd)).max() < 1e-12


def test_parametrised_gates_registers_and_broadcast():
    src = HEAD + """qreg a[2]; qreg b[1];
    h a; rx(pi/2) b[0]; u3(0.1,0.2,0.3) a[1]; rz(-pi/4) a[0]; cu1(pi/8) a[0],b[0]; crz(0.3) b[0],a[1];
    swap a[0],a[1]; sdg a[0]; tdg b[0]; rzz(0.4) a[1],b[0]; cz a[0],b[0]; p(0.7) a[1]; u2(0.1,0.2) b[0];"""
    n, ops = qasm_to_ops(src)
    assert n == 3 and [q for q, _ in ops[:2]] == [[0], [1]]          # h broadcast over register a
    psi = state(n, ops)
    assert abs(np.vdot(psi, psi).real - 1) < 1e-12
    for _, U in ops:
        assert np.abs(U.conj().T @ U - np.eye(len(U))).max() < 1e-12


def test_toffoli_cswap_and_user_gates():
    src = HEAD + """qreg q[3];
    gate majority a,b,c { cx c,b; cx c,a; ccx a,b,c; }
    gate myrot(t) x { ry(t/2) x; ry(t/2) x; }
    x q[0]; x q[1]; ccx q[0],q[1],q[2]; myrot(pi) q[0]; majority q[0],q[1],q[2]; cswap q[2],q[0],q[1];"""
    n, ops = qasm_to_ops(src)
    psi = state(n, ops)
    # reference evaluation with explicit permutation / rotation matrices
    ref = np.zeros(8, dtype=np.complex128)
    ref[0] = 1

    def perm(f):
        nonlocal ref
        new = np.zeros_like(ref)
        for i in range(8):
            new[f(i)] += ref[i]
        ref = new

    bit = lambda i, q: (i >> q) & 1                                   # noqa: E731
    perm(lambda i: i ^ 1); perm(lambda i: i ^ 2)
    perm(lambda i: i ^ (4 if bit(i, 0) and bit(i, 1) else 0))
    ry = np.array([[math.cos(math.pi / 2), -math.sin(math.pi / 2)], [math.sin(math.pi / 2), math.cos(math.pi / 2)]])
    O.apply_1q(ref, 0, ry.astype(np.complex128))
    perm(lambda i: i ^ (2 if bit(i, 2) else 0)); perm(lambda i: i ^ (1 if bit(i, 2) else 0))
    perm(lambda i: i ^ (4 if bit(i, 0) and bit(i, 1) else 0))
    perm(lambda i: (i & 4) | (((i >> 1) & 1) | ((i & 1) << 1)) if bit(i, 2) else i)
    assert np.abs(psi - ref).max() < 1e-12


def test_errors():
    with pytest.raises(QasmError, match="unsupported gate"):
        qasm_to_ops(HEAD + "qreg q[1]; foo q[0];")
    with pytest.raises(QasmError, match="not supported"):
        qasm_to_ops(HEAD + "qreg q[1]; h q[0]; reset q[0];")
    with pytest.raises(QasmError, match="out of range"):
        qasm_to_ops(HEAD + "qreg q[1]; h q[3];")
    with pytest.raises(QasmError, match="no name in the reference"):
        qasm_to_dict(HEAD + "qreg q[1]; rz(0.1) q[0];")


def test_parameter_expressions_are_walked_not_evaluated():
    """A .qasm file is untrusted input: attribute chains, subscripts, lambdas, foreign calls are refused
    (they used to reach interpreter internals through eval()); the arithmetic of the spec still works."""
    import math
    from quantum_simulations_b200.circuit.qasm import _eval
    for bad in ("().__class__.__base__.__subclasses__().__len__()", "pi.real", "[1][0]", "(lambda: 1)()",
                "__import__('os')", "open('x')", "1 if 1 else 2", "1 < 2", "sin(1, 2)", "sin.__name__", "foo", "2 ** 99999",
                "'a'", "True", "1j"):
        with pytest.raises(QasmError):
            _eval(bad, {})
        with pytest.raises(QasmError):
            qasm_to_ops(HEAD + f"qreg q[1]; rz({bad}) q[0];")
    assert _eval("-pi/2 + 2^3*sin(pi/6) - sqrt(4)*ln(exp(1))", {}) == pytest.approx(-math.pi / 2 + 8 * 0.5 - 2.0)
    assert _eval("theta/2 + 1e-3", {"theta": 0.5}) == pytest.approx(0.251)
    with pytest.raises(QasmError, match="cannot evaluate"):
        _eval("1/0", {})


def test_reset_before_any_gate_is_the_identity_and_later_resets_are_refused():
    from quantum_simulations_b200.circuit.qasm import QasmError, qasm_to_ops
    head = 'OPENQASM 2.0;\ninclude "qelib1.inc";\nqreg q[3];\ncreg c[3];\n'
    n, ops = qasm_to_ops(head + "reset q[0];\nreset q;\nh q[0];\ncx q[0],q[1];\nreset q[2];\n")
    assert n == 3 and [qs for qs, _ in ops] == [[0], [0, 1]]
    with pytest.raises(QasmError, match="after a gate on the same qubit"):
        qasm_to_ops(head + "h q[1];\nreset q[1];\n")
    with pytest.raises(QasmError, match="not a unitary gate"):
        qasm_to_ops(head + "h q[1];\nmeasure q[1] -> c[1];\nif(c==1) x q[0];\n")


@pytest.mark.parametrize("name", ["rxx", "ryy"])
def test_rxx_ryy_lower_to_one_qubit_gates_around_a_diagonal(name):
    """exp(-i t/2 P(x)P) arrives as 1-qubit gates around RZZ (diagonal), so that it plans into fused
    passes instead of a dense 4x4 sweep; the product must be the same unitary."""
    import math
    from oracle import ref_dense as O
    from quantum_simulations_b200.circuit.passes import PassCompiler
    P = {"rxx": np.array([[0, 1], [1, 0]], complex), "ryy": np.array([[0, -1j], [1j, 0]])}[name]
    t = 0.613
    n, ops = qasm_to_ops(HEAD + f"qreg q[3]; h q[0]; ry(0.4) q[2]; {name}({t}) q[2],q[0];")
    assert all(len(qs) == 1 or np.count_nonzero(U - np.diag(np.diag(U))) == 0 for qs, U in ops)
    psi = np.zeros(8, complex); psi[0] = 1
    for qs, U in ops:
        O.apply_1q(psi, qs[0], U) if len(qs) == 1 else O.apply_2q(psi, qs[0], qs[1], U)
    phi = np.zeros(8, complex); phi[0] = 1
    for qs, U in ops[:2]:
        O.apply_1q(phi, qs[0], U)
    O.apply_2q(phi, 2, 0, math.cos(t / 2) * np.eye(4) - 1j * math.sin(t / 2) * np.kron(P, P))
    assert np.abs(psi - phi).max() <= 1e-14
    prog = PassCompiler(12, tile_bits=7, low_bits=2).compile(qasm_to_ops(HEAD + f"qreg q[12]; h q; {name}(0.3) q[11],q[1];")[1])
    assert prog.stats["dense2q_steps"] == 0


def _random_qasm(seed: int):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(5, 11))
    src = ["OPENQASM 2.0;", 'include "qelib1.inc";', f"qreg q[{n}];"]
    for _ in range(int(rng.integers(10, 50))):
        k = rng.random()
        a, b = (int(x) for x in rng.choice(n, 2, replace=False))
        t = float(rng.normal())
        if k < 0.2:
            src.append(f"{['h', 'x', 't', 's', 'sdg', 'tdg', 'z', 'y'][int(rng.integers(8))]} q[{a}];")
        elif k < 0.35:
            src.append(f"{['rx', 'ry', 'rz', 'u1'][int(rng.integers(4))]}({t}) q[{a}];")
        elif k < 0.45:
            src.append(f"u3({t},{t / 2},{-t}) q[{a}];")
        elif k < 0.6:                                                     # compiled ZZ
            src += [f"cx q[{a}],q[{b}];", f"rz({t}) q[{b}];", f"cx q[{a}],q[{b}];"]
        elif k < 0.7:                                                     # compiled controlled phase
            src += [f"u1({t / 2}) q[{a}];", f"cx q[{a}],q[{b}];", f"u1({-t / 2}) q[{b}];", f"cx q[{a}],q[{b}];", f"u1({t / 2}) q[{b}];"]
        elif k < 0.78:
            src.append(f"{['rxx', 'ryy', 'rzz'][int(rng.integers(3))]}({t}) q[{a}],q[{b}];")
        elif k < 0.86:
            src.append(f"{['cx', 'cz', 'cy', 'ch', 'swap'][int(rng.integers(5))]} q[{a}],q[{b}];")
        elif k < 0.92:
            src.append(f"{['cu1', 'crz', 'cp'][int(rng.integers(3))]}({t}) q[{a}],q[{b}];")
        else:
            c = int(rng.choice([x for x in range(n) if x not in (a, b)]))
            src.append(f"{['ccx', 'cswap'][int(rng.integers(2))]} q[{a}],q[{b}],q[{c}];")
    return n, "\n".join(src)


def test_random_programs_through_fusion_and_the_planner():
    """Seeded random OpenQASM programs (6,000 seeds of this generator ran clean when it was written):
    front end -> fuse_2q_blocks(tol=1e-14) -> plan_single -> pass emulator, against the gate-by-gate oracle."""
    from quantum_simulations_b200.circuit import sharding
    from quantum_simulations_b200.circuit.fusion import fuse_2q_blocks
    from tests.pass_emulator import run_program
    fused_away = 0
    for seed in range(120):
        n, text = _random_qasm(seed)
        _, ops = qasm_to_ops(text)
        want = np.zeros(1 << n, dtype=np.complex128)
        want[0] = 1
        O.apply_ops(want, ops)
        fo = fuse_2q_blocks(ops, tol=1e-14)
        fused_away += len(ops) - len(fo)
        prog = sharding.plan_single(fo, n, "complex128", True, False, tile_bits=min(n, 4 + seed % 4), low_bits=seed % 3)
        psi = np.random.default_rng(seed).standard_normal(1 << n) + 0j if prog.fused_init else None
        if psi is None:
            psi = np.zeros(1 << n, dtype=np.complex128)
            psi[0] = 1
        psi = run_program(prog, psi)
        assert np.abs(psi - want).max() <= 1e-11, seed
    assert fused_away > 1000


TELEPORT = HEAD + """qreg q[3]; creg c0[1]; creg c1[1];
ry(0.9) q[0]; t q[0];
h q[1]; cx q[1],q[2];
cx q[0],q[1]; h q[0];
measure q[0] -> c0[0]; measure q[1] -> c1[0];
if(c1==1) x q[2];
if(c0==1) z q[2];
reset q[0]; reset q[1];
"""


def test_steps_keep_measure_reset_and_if():
    from quantum_simulations_b200.circuit.qasm import qasm_to_steps, is_unitary_program
    n, steps, cregs = qasm_to_steps(TELEPORT)
    assert n == 3 and cregs == {"c0": 1, "c1": 1}
    kinds = [s[0] for s in steps]
    assert kinds == ["ops", "measure", "measure", "if", "if", "reset", "reset"]
    assert steps[1][1:] == (0, "c0", 0) and steps[3][1:3] == ("c1", 1) and [qs for qs, _ in steps[3][3]] == [[2]]
    assert not is_unitary_program(steps)
    # a program whose measurements are only read at the end stays one unitary circuit
    _, s2, _ = qasm_to_steps(HEAD + "qreg q[2]; creg c[2]; h q[0]; cx q[0],q[1]; measure q -> c;")
    assert [s[0] for s in s2] == ["ops", "measure", "measure"] and is_unitary_program(s2)
    _, s3, _ = qasm_to_steps(HEAD + "qreg q[2]; creg c[2]; h q[0]; measure q[0] -> c[0]; h q[0];")
    assert not is_unitary_program(s3)
    with pytest.raises(QasmError, match="unknown classical register"):
        qasm_to_steps(HEAD + "qreg q[1]; if(c==1) x q[0];")
    with pytest.raises(QasmError, match="only gate statements"):
        qasm_to_steps(HEAD + "qreg q[1]; creg c[1]; if(c==1) reset q[0];")


@pytest.mark.parametrize("seed", range(8))
def test_teleportation_trajectory_known_answer(seed):
    """Whatever the two measurement outcomes are, qubit 2 ends in T RY(0.9)|0> and qubits 0, 1 in |00>
    (oracle definition of one trajectory, oracle/ref_dense.py::run_qasm_steps)."""
    import math
    from quantum_simulations_b200.circuit.qasm import qasm_to_steps
    n, steps, cregs = qasm_to_steps(TELEPORT)
    psi, bits = O.run_qasm_steps(n, steps, cregs, seed)
    a, b = math.cos(0.45), math.sin(0.45) * np.exp(0.25j * math.pi)
    want = np.zeros(8, dtype=np.complex128)
    want[0], want[4] = a, b
    assert abs(np.vdot(psi, psi).real - 1) < 1e-12
    assert abs(abs(np.vdot(want, psi)) - 1) < 1e-12          # equal up to a global phase
    assert set(bits) == {"c0", "c1"} and all(v in (0, 1) for v in bits.values())
