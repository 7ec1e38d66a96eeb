"""The pass compiler against the oracle, through the NumPy pass emulator (no GPU needed)."""
import numpy as np
import pytest

from oracle import ref_dense as O
from quantum_simulations_b200.circuit.io import validate_circuit_dict
from quantum_simulations_b200.circuit.passes import PassCompiler, lower_op, MicroOp, Dense2Q
from quantum_simulations_b200.kernel import gates as G
from quantum_simulations_b200 import workloads as W
from quantum_simulations_b200 import _lib as L
from tests.pass_emulator import run_program


def ir_ops(cd):
    cd = validate_circuit_dict(cd)
    return [(g["qubits"], G.gate_matrix(g["gate"], g["params"])) for g in cd["gates"]]


def check(cd, **kw):
    n = cd["number_of_qubits"]
    prog = PassCompiler(n, **kw).compile(ir_ops(cd))
    psi = np.zeros(1 << n, dtype=np.complex128)
    psi[0] = 1
    run_program(prog, psi)
    want = O.simulate(validate_circuit_dict(cd))
    assert prog.final_pos == list(range(n))
    assert np.abs(psi - want).max() <= 1e-12
    return prog


@pytest.mark.parametrize("n,t,a", [(6, 6, 5), (7, 5, 1), (8, 6, 2), (9, 7, 3), (10, 8, 3), (10, 12, 5), (12, 8, 4)])
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_mixed(n, t, a, seed):
    check(W.random_mixed(n, 120, seed), tile_bits=t, low_bits=a)


@pytest.mark.parametrize("n,t,a", [(8, 6, 2), (10, 7, 3), (12, 8, 3), (12, 12, 5)])
def test_random_1q_cz(n, t, a):
    check(W.random_1q_cz(n, 20, 1234), tile_bits=t, low_bits=a)


@pytest.mark.parametrize("n", [5, 8, 11])
def test_qft_and_ghz(n):
    check(W.qft(n), tile_bits=min(n, 7), low_bits=2)
    check(W.ghz(n), tile_bits=min(n, 6), low_bits=2)


@pytest.mark.parametrize("max_rounds", [2, 3, 6])
def test_round_budget(max_rounds):
    prog = check(W.random_1q_cz(10, 12, 7), tile_bits=8, low_bits=3, max_rounds=max_rounds)
    assert all(s.desc.n_rounds <= max_rounds + 2 for s in prog.passes)


def test_reference_staging_bug_case_is_right_here():
    """H(2) H(0) CZ(2,0) H(0) H(1) — the reference's heuristic stager
    reorders H0·CZ·H0 into H0·H0·CZ.  Per-qubit program order must be kept."""
    cd = {"number_of_qubits": 5, "gates": [
        {"qubits": [2], "gate": "H"}, {"qubits": [0], "gate": "H"}, {"qubits": [2, 0], "gate": "CZ"},
        {"qubits": [0], "gate": "H"}, {"qubits": [1], "gate": "H"}]}
    check(cd, tile_bits=4, low_bits=0)


def test_swap_is_a_rename_not_arithmetic():
    cd = {"number_of_qubits": 6, "gates": [
        {"qubits": [0], "gate": "H"}, {"qubits": [0, 5], "gate": "SWAP"}, {"qubits": [5], "gate": "T"},
        {"qubits": [0], "gate": "X"}, {"qubits": [1, 4], "gate": "SWAP"}, {"qubits": [4, 0], "gate": "CNOT"}]}
    prog = check(cd, tile_bits=5, low_bits=1)
    assert all(s.n_micro_ops <= 5 for s in prog.passes)   # 4 gates + 1 SCALE


def test_dense_2q_runs_as_own_step():
    rng = np.random.default_rng(5)
    q, _ = np.linalg.qr(rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4)))
    n = 7
    ops = [([0], G.H()), ([3], G.H()), ([3, 1], q), ([1], G.T()), ([6, 3], G.CNOT())]
    prog = PassCompiler(n, tile_bits=5, low_bits=1).compile(ops)
    psi = np.zeros(1 << n, dtype=np.complex128); psi[0] = 1
    run_program(prog, psi)
    want = np.zeros(1 << n, dtype=np.complex128); want[0] = 1
    O.apply_ops(want, ops)
    assert prog.stats["dense2q_steps"] == 1
    assert np.abs(psi - want).max() <= 1e-12


def test_lowering_structures():
    kinds = lambda ops: [o.kind for o in ops]
    assert kinds(lower_op([3], G.H())) == [L.OP_HAD, L.OP_SCALE]
    assert kinds(lower_op([3], G.Y())) == [L.OP_YSWAP]
    assert kinds(lower_op([3], G.X())) == [L.OP_XSWAP]
    assert kinds(lower_op([3], G.RY(0.3))) == [L.OP_ROT]
    assert kinds(lower_op([3], G.RY(4.0))) == [L.OP_ROT, L.OP_SIGN]        # cos < 0: global -1
    assert [(o.kind, o.ctrls) for o in lower_op([3], G.Z())] == [(L.OP_SIGN, (3,))]
    assert [(o.kind, o.ctrls) for o in lower_op([3], G.T())] == [(L.OP_PHASE, (3,))]
    assert [(o.kind, o.ctrls) for o in lower_op([3], G.S())] == [(L.OP_PHASE, (3,))]
    assert [(o.kind, o.ctrls) for o in lower_op([2, 5], G.CZ())] == [(L.OP_SIGN, (2, 5))]
    assert [(o.kind, o.target, o.ctrls) for o in lower_op([2, 5], G.CNOT())] == [(L.OP_XSWAP, 5, (2,))]
    assert [(o.kind, o.target, o.ctrls) for o in lower_op([2, 5], G.CY())] == [(L.OP_YSWAP, 5, (2,))]
    assert lower_op([2, 5], G.SWAP()) == [("swap", 2, 5)]
    assert lower_op([1], np.eye(2)) == []
    # a general unitary becomes phase . rotation . phase (. global phase)
    assert set(kinds(lower_op([3], G.T() @ G.H()))) <= {L.OP_PHASE, L.OP_ROT, L.OP_SIGN}
    # control on qubits[1]; controlled-H is a controlled sign + rotation
    cu = np.eye(4, dtype=complex); cu[np.ix_([1, 3], [1, 3])] = G.H()
    low = lower_op([2, 5], cu)
    assert all(5 in o.ctrls for o in low) and L.OP_ROT in kinds(low)
    # every PHASE / ROT coefficient stays in the stable range |t| <= 1
    for seed in range(20):
        rng = np.random.default_rng(seed)
        q, r = np.linalg.qr(rng.standard_normal((2, 2)) + 1j * rng.standard_normal((2, 2)))
        for o in lower_op([0], q):
            if o.kind in (L.OP_PHASE, L.OP_ROT):
                assert abs(o.m[0]) <= 1 + 1e-12
    # non-unitary blocks cannot be lowered: they run through the dense per-gate kernels
    from quantum_simulations_b200.circuit.passes import Dense1Q
    assert isinstance(lower_op([0], np.array([[1, 2], [3, 4]]))[0], Dense1Q)
    assert isinstance(lower_op([0, 1], np.diag([1, 2, 1, 1]))[0], Dense2Q)


def test_arbitrary_unitaries_and_non_unitary_fallback():
    rng = np.random.default_rng(3)
    n = 7
    ops = []
    for k in range(40):
        q, r = np.linalg.qr(rng.standard_normal((2, 2)) + 1j * rng.standard_normal((2, 2)))
        tq = int(rng.integers(n))
        if k % 3 == 0:
            c = int((tq + 1 + rng.integers(n - 1)) % n)
            cu = np.eye(4, dtype=complex); cu[2:, 2:] = q
            ops.append(([c, tq], cu))
        else:
            ops.append(([tq], q))
    ops.insert(10, ([2], np.array([[1, 0.5], [0.25, 1]], dtype=complex)))      # non-unitary
    prog = PassCompiler(n, tile_bits=5, low_bits=1).compile(ops)
    psi = np.zeros(1 << n, dtype=np.complex128); psi[0] = 1
    run_program(prog, psi)
    want = np.zeros(1 << n, dtype=np.complex128); want[0] = 1
    O.apply_ops(want, ops)
    assert prog.stats["dense1q_steps"] == 1
    assert np.abs(psi - want).max() <= 1e-12


def test_rank_bit_controls_and_nonlocal_targets():
    """n=8 state split over 4 shards (n_local=6): diagonal / control use of rank bits is free,
    mixing a rank bit raises like the reference's local kernels do (cpu_scalar.py:13-18)."""
    n, n_local = 8, 6
    ops = [([0], G.H()), ([7], G.Z()), ([7, 2], G.CNOT()), ([6, 7], G.CZ()), ([1, 6], G.CR(3))]
    for rank in range(4):
        full = np.zeros(1 << n, dtype=np.complex128)
        rng = np.random.default_rng(rank)
        full[:] = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
        want = full.copy()
        O.apply_ops(want, ops)
        prog = PassCompiler(n, n_local, tile_bits=5, low_bits=1).compile(ops)
        shard = full[rank << n_local:(rank + 1) << n_local].copy()
        run_program(prog, shard, rank=rank)
        assert np.abs(shard - want[rank << n_local:(rank + 1) << n_local]).max() <= 1e-12
    with pytest.raises(NotImplementedError, match="non-local"):
        PassCompiler(n, n_local, tile_bits=5, low_bits=1, allow_swaps=False).compile([([7], G.H())])
    prog = PassCompiler(n, n_local, tile_bits=5, low_bits=1).compile([([7], G.H())])   # default: swap it in
    assert prog.stats["swaps"] == 2


def test_headline_plan_sizes():
    """Planning the BASELINE.json circuits is cheap and collapses layers into few passes."""
    prog = PassCompiler(28).compile(ir_ops(W.qft(28)))
    assert prog.stats["passes"] <= 6          # 55 levels, 406 gates
    prog = PassCompiler(30).compile(ir_ops(W.random_1q_cz(30, 20, 1234)))
    assert prog.stats["passes"] <= 16         # 20 levels, 445 gates


def test_x_and_y_gates_never_touch_data():
    """The Pauli-X frame: uncontrolled X / Y gates become address flips of the final pass."""
    cd = W.random_1q_cz(12, 20, 1234)
    prog = check(cd, tile_bits=8, low_bits=3)
    kinds = [s.ops[i].kind for s in prog.passes for i in range(s.n_micro_ops)]
    assert L.OP_XSWAP not in kinds and L.OP_YSWAP not in kinds
    assert any(s.desc.store_flip for s in prog.passes)
    assert prog.final_flips == [0] * 12
    # controlled X survives as a real op (GHZ), with the frame applied to its control
    cd = {"number_of_qubits": 6, "gates": [
        {"qubits": [0], "gate": "X"}, {"qubits": [0, 1], "gate": "CNOT"}, {"qubits": [1], "gate": "Y"},
        {"qubits": [1, 2], "gate": "CY"}, {"qubits": [2], "gate": "H"}, {"qubits": [2, 3], "gate": "CZ"},
        {"qubits": [0, 3], "gate": "CR", "params": {"k": 3}}, {"qubits": [3], "gate": "RY", "params": {"theta": 0.7}},
        {"qubits": [3], "gate": "X"}, {"qubits": [3], "gate": "RY", "params": {"theta": 1.9}},
        {"qubits": [3, 4], "gate": "CU", "params": {"U": [[0.6, -0.8], [0.8, 0.6]], "exponent": 1}}]}
    check(cd, tile_bits=4, low_bits=0)
    check(cd, tile_bits=6, low_bits=2, x_frame=False)


@pytest.mark.parametrize("n,t,a", [(9, 6, 2), (11, 7, 3), (12, 8, 3)])
def test_zero_support_skipping_is_exact(n, t, a):
    """compile(zero_state=True): early passes name only the tiles that can hold data; the emulator
    asserts that every skipped tile is exactly zero, and the result equals the oracle's."""
    for cd in (W.random_1q_cz(n, 20, 1234), W.qft(n), W.ghz(n), W.random_mixed(n, 150, 4)):
        ops = ir_ops(cd)
        prog = PassCompiler(n, tile_bits=t, low_bits=a).compile(ops, zero_state=True)
        assert prog.passes[0].desc.n_active == 0                     # first pass: one tile
        assert any(s.desc.n_active == -1 for s in prog.passes) or n - t <= 3
        psi = np.zeros(1 << n, dtype=np.complex128)
        psi[0] = 1
        run_program(prog, psi)
        assert np.abs(psi - O.simulate(validate_circuit_dict(cd))).max() <= 1e-12


def test_fused_initialisation_ignores_the_input():
    """plan_single / sharding.plan mark the first pass zero_input: the program starts from |0...0>
    whatever the shard holds (no memset, no read in the first pass)."""
    from quantum_simulations_b200.circuit import sharding
    from tests.pass_emulator import run_program_sharded
    n = 11
    cd = W.random_1q_cz(n, 20, 1234)
    ops = ir_ops(cd)
    want = O.simulate(validate_circuit_dict(cd))
    prog = sharding.plan_single(ops, n, tile_bits=7, low_bits=2)
    assert prog.fused_init and prog.passes[0].desc.zero_input == 1 and all(s.desc.zero_input == 0 for s in prog.passes[1:])
    junk = np.random.default_rng(0).standard_normal(1 << n) + 1j
    run_program(prog, junk)
    assert np.abs(junk - want).max() <= 1e-12
    prog = sharding.plan(ops, n, n - 2, tile_bits=7, low_bits=2)
    assert prog.fused_init
    junk = run_program_sharded(prog, np.random.default_rng(1).standard_normal(1 << n) + 1j)
    if prog.rank_flip_mask:
        sh = junk.reshape(4, -1)
        junk = np.concatenate([sh[r ^ prog.rank_flip_mask] for r in range(4)])
    assert np.abs(junk - want).max() <= 1e-12
    assert not PassCompiler(n, tile_bits=7, low_bits=2).compile(ops).fused_init          # plain compile: never


def test_lowering_cache_follows_the_contents_of_the_op_list():
    """One compiler, the SAME list object mutated in place between two compile() calls (a parameter sweep):
    the second program must be the second circuit's (the cache used to be keyed on id() and length)."""
    n = 9
    a = ir_ops(W.random_1q_cz(n, 8, 5))
    b = ir_ops(W.random_1q_cz(n, 8, 6))
    assert len(a) == len(b)
    comp = PassCompiler(n, tile_bits=6, low_bits=2)
    ops = list(a)
    for want_cd in (W.random_1q_cz(n, 8, 5), W.random_1q_cz(n, 8, 6)):
        prog = comp.compile(ops)
        psi = np.zeros(1 << n, dtype=np.complex128)
        psi[0] = 1
        run_program(prog, psi)
        assert np.abs(psi - O.simulate(validate_circuit_dict(want_cd))).max() <= 1e-12
        ops[:] = b
