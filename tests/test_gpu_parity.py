"""GPU parity: the CUDA path (through the C ABI) against the oracle and the reference's
golden vectors.  Tolerances are the north star's: max|Δamp| <= 1e-12 (complex128),
<= 1e-5 (complex64).  Nothing here reads /root/reference."""
import tempfile

import numpy as np
import pytest

from oracle import ref_dense as O
from oracle import c_oracle as CO
from quantum_simulations_b200 import workloads as W
from quantum_simulations_b200.circuit.io import validate_circuit_dict
from quantum_simulations_b200.kernel import gates as G
from tests._specs import STATE_SPECS, RUNNER_SPECS, circuit_from_spec

pytestmark = pytest.mark.gpu

TOL = {"complex128": 1e-12, "complex64": 1e-5}


def _cuda():
    from quantum_simulations_b200.kernel import cuda
    return cuda


def _rand_state(n, seed, dtype="complex128"):
    rng = np.random.default_rng(seed)
    v = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    v /= np.linalg.norm(v)
    return v.astype(dtype)


def _rand_unitary(dim, seed):
    rng = np.random.default_rng(seed)
    q, r = np.linalg.qr(rng.standard_normal((dim, dim)) + 1j * rng.standard_normal((dim, dim)))
    return q * (np.diag(r) / np.abs(np.diag(r)))


# ------------------------------------------------------------------ per-gate kernels
@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("n", [1, 2, 5, 9, 14])
def test_apply_1q_every_qubit(n, dtype):
    cuda = _cuda()
    psi = _rand_state(n, n, dtype)
    want = psi.astype(np.complex128)
    with cuda.DeviceState(n, dtype) as st:
        st.upload(psi)
        for q in range(n):
            U = _rand_unitary(2, 10 * n + q)
            st.apply_1q(q, U)
            O.apply_1q(want, q, U)
        got = st.download()
    assert got.dtype == np.dtype(dtype)
    assert np.abs(got - want).max() <= TOL[dtype]


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_apply_2q_all_pairs(dtype):
    cuda = _cuda()
    n = 7
    psi = _rand_state(n, 3, dtype)
    want = psi.astype(np.complex128)
    with cuda.DeviceState(n, dtype) as st:
        st.upload(psi)
        for qa in range(n):
            for qb in range(n):
                if qa != qb:
                    U = _rand_unitary(4, 100 * qa + qb)
                    st.apply_2q(qa, qb, U)
                    O.apply_2q(want, qa, qb, U)
        got = st.download()
    assert np.abs(got - want).max() <= TOL[dtype] * 4


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_diag_ctrl_kq(dtype):
    cuda = _cuda()
    n = 9
    psi = _rand_state(n, 4, dtype)
    want = psi.astype(np.complex128)
    rng = np.random.default_rng(8)
    with cuda.DeviceState(n, dtype) as st:
        st.upload(psi)
        # diagonal on 1..4 qubits
        for qs in ([3], [0, 8], [5, 2, 7], [1, 4, 6, 0]):
            ph = np.exp(1j * rng.uniform(0, 2 * np.pi, 1 << len(qs)))
            st.apply_diag(qs, ph)
            idx = np.arange(1 << n)
            t = np.zeros_like(idx)
            for q in qs:
                t = (t << 1) | ((idx >> q) & 1)
            want *= ph[t]
        # controlled 1q
        for c, t in ((0, 5), (8, 1), (3, 2)):
            U = _rand_unitary(2, c * 10 + t)
            st.apply_ctrl_1q(c, t, U)
            cu = np.eye(4, dtype=complex); cu[2:, 2:] = U
            O.apply_2q(want, c, t, cu)
        # dense k-qubit
        for qs in ([4], [6, 1], [0, 7, 3], [8, 2, 5, 1], [3, 0, 6, 8, 4]):
            k = len(qs)
            U = _rand_unitary(1 << k, 77 + k)
            st.apply_kq(qs, U)
            # reference semantics: row bit (k-1-i) <-> qs[i]
            idx = np.arange(1 << n)
            mask = sum(1 << q for q in qs)
            base = idx[(idx & mask) == 0]
            offs = [sum(((r >> (k - 1 - i)) & 1) << qs[i] for i in range(k)) for r in range(1 << k)]
            v = np.stack([want[base | o] for o in offs])
            out = U @ v
            for r, o in enumerate(offs):
                want[base | o] = out[r]
        got = st.download()
    assert np.abs(got - want).max() <= TOL[dtype] * 8


def test_apply_op_dispatch_all_gates():
    cuda = _cuda()
    n = 6
    cd = validate_circuit_dict(W.random_mixed(n, 200, 99))
    with cuda.DeviceState(n) as st:
        st.init_zero()
        for g in cd["gates"]:
            st.apply_op(g["qubits"], G.gate_matrix(g["gate"], g["params"]))
        got = st.download()
    assert np.abs(got - O.simulate(cd)).max() <= 1e-12


def test_a1_a2_host_callables_and_nonlocal_raise():
    cuda = _cuda()
    chunk = _rand_state(6, 1)
    want = chunk.copy()
    U1, U2 = _rand_unitary(2, 1), _rand_unitary(4, 2)
    cuda.apply_1q(chunk, 3, U1); O.apply_1q(want, 3, U1)
    cuda.apply_2q(chunk, 5, 0, U2); O.apply_2q(want, 5, 0, U2)
    assert np.abs(chunk - want).max() <= 1e-12
    small = np.zeros(4, dtype=np.complex128); small[0] = 1
    with pytest.raises(NotImplementedError, match="non-local"):
        cuda.apply_1q(small, 2, G.H())
    with pytest.raises(NotImplementedError, match="non-local"):
        cuda.apply_2q(small, 0, 2, G.CNOT())


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_nonlocal_butterfly_drop_ins_match_oracle(dtype):
    """kernel.cuda_nonlocal mirrors cpu_nonlocal.py:22-67 (same names, in place on host chunks)."""
    from quantum_simulations_b200.kernel import cuda_nonlocal as KN
    rng = np.random.default_rng(11)
    tol = TOL[dtype] * 10

    def chunks(m):
        return [(rng.standard_normal(256) + 1j * rng.standard_normal(256)).astype(dtype) for _ in range(m)]

    U1 = G.gate_matrix("RY", {"theta": 0.7}) @ G.H()
    U2 = np.kron(G.H(), G.gate_matrix("RY", {"theta": 1.1})) @ G.CNOT()
    for fn, m, args in ((KN.apply_1q_pair, 2, (U1,)), (KN.apply_2q_pair_qa_local, 2, (3, U2)),
                        (KN.apply_2q_pair_qb_local, 2, (2, U2)), (KN.apply_2q_quad, 4, (U2,))):
        got = chunks(m)
        want = [c.astype(np.complex128) for c in got]
        getattr(O, fn.__name__)(*want, *args)
        fn(*got, *args)
        for a, b in zip(got, want):
            assert a.dtype == np.dtype(dtype) and np.abs(a - b).max() <= tol, fn.__name__


def test_sharded_handle_semantics_on_one_gpu():
    """4 shards of an 8-qubit state as 4 handles on one device: rank bits as controls /
    diagonal qubits work without communication; mixing one raises QSV_ENONLOCAL."""
    cuda = _cuda()
    n, world = 8, 4
    n_local = 6
    full = _rand_state(n, 21)
    want = full.copy()
    ops = [([0], G.H()), ([7], G.Z()), ([7, 2], G.CNOT()), ([6, 7], G.CZ()), ([1, 6], G.CR(3))]
    O.apply_ops(want, ops)
    for rank in range(world):
        with cuda.DeviceState(n, rank=rank, world=world) as st:
            st.upload(full[rank << n_local:(rank + 1) << n_local])
            for qs, U in ops:
                st.apply_op(qs, U)
            got = st.download()
            with pytest.raises(NotImplementedError, match="non-local"):
                st.apply_1q(7, G.H())
        assert np.abs(got - want[rank << n_local:(rank + 1) << n_local]).max() <= 1e-12


# ------------------------------------------------------------------ fused pass path
@pytest.mark.parametrize("spec", STATE_SPECS)
def test_golden_states_fused_and_unfused(golden, spec):
    from quantum_simulations_b200.kernel.cuda_dense import simulate
    want = golden[f"state/{spec}"]
    cd = circuit_from_spec(spec)
    for fused in (False, True):
        got = simulate(cd, fused=fused)
        assert np.abs(got - want).max() <= 1e-12, (spec, fused)
    got64 = simulate(cd, dtype="complex64")
    assert got64.dtype == np.complex64 and np.abs(got64 - want).max() <= 1e-5


@pytest.mark.parametrize("n,t,a", [(10, 6, 2), (12, 8, 3), (14, 10, 4), (14, 12, 5), (16, 12, 5),
                                   (16, 11, 3), (18, 12, 6)])
def test_pass_kernel_tile_shapes(n, t, a):
    from quantum_simulations_b200.kernel.cuda_dense import simulate
    cd = W.random_mixed(n, 300, n * 7 + t)
    want = CO.simulate_c(validate_circuit_dict(cd))
    got = simulate(cd, tile_bits=t, low_bits=a)
    assert np.abs(got - want).max() <= 1e-12


@pytest.mark.parametrize("dtype,t", [("complex64", 13), ("complex64", 9), ("complex128", 12)])
def test_random_1q_cz_depth20(dtype, t):
    from quantum_simulations_b200.kernel.cuda_dense import simulate
    n = 20
    cd = W.random_1q_cz(n, 20, 1234)
    want = CO.simulate_c(validate_circuit_dict(cd))
    got = simulate(cd, dtype=dtype, tile_bits=t)
    assert np.abs(got - want).max() <= TOL[dtype]


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("jit", [True, False])
@pytest.mark.parametrize("workload", ["random_1q_cz", "random_mixed", "qft", "ghz"])
def test_specialised_and_interpreted_passes_agree_with_oracle(workload, jit, dtype):
    """The run-time specialised kernels (csrc/jit.cuh) and the interpreting ring kernel run the
    same compiled passes; both must match the C oracle (default ring tiles: 2^11 amplitudes)."""
    from quantum_simulations_b200.kernel.cuda_dense import simulate
    n = 18
    cd = {"random_1q_cz": lambda: W.random_1q_cz(n, 20, 1234), "random_mixed": lambda: W.random_mixed(n, 400, 5),
          "qft": lambda: W.qft(n), "ghz": lambda: W.ghz(n)}[workload]()
    want = CO.simulate_c(validate_circuit_dict(cd))
    got = simulate(cd, jit=jit, dtype=dtype)
    assert got.dtype == np.dtype(dtype) and np.abs(got - want).max() <= TOL[dtype]


@pytest.mark.parametrize("workload", ["random_1q_cz", "random_mixed", "qft", "ghz"])
def test_zero_support_skipping_matches_oracle(workload):
    """simulate(skip_zero_support=True): early passes visit only the tiles that can hold data."""
    from quantum_simulations_b200.kernel.cuda_dense import simulate
    n = 20
    cd = {"random_1q_cz": lambda: W.random_1q_cz(n, 20, 1234), "random_mixed": lambda: W.random_mixed(n, 400, 5),
          "qft": lambda: W.qft(n), "ghz": lambda: W.ghz(n)}[workload]()
    want = CO.simulate_c(validate_circuit_dict(cd))
    for dtype in ("complex128", "complex64"):
        got = simulate(cd, dtype=dtype, skip_zero_support=True)
        assert np.abs(got - want).max() <= TOL[dtype]


@pytest.mark.parametrize("jit", [True, False])
def test_fused_initialisation_on_the_device(jit):
    """A program whose first pass has zero_input starts from |0...0> whatever the shard holds —
    on the specialised kernels (no read at all) and on the interpreting ones (they initialise first)."""
    from quantum_simulations_b200.kernel.cuda import DeviceState
    from quantum_simulations_b200.kernel.cuda_dense import compile_circuit
    n = 18
    cd = W.random_1q_cz(n, 20, 1234)
    prog = compile_circuit(cd)
    assert prog.fused_init
    with DeviceState(n) as st:
        st.upload(np.full(1 << n, 0.5 + 0.25j))            # junk that must be ignored
        st.run_program(prog, jit=jit)
        got = st.download()
    assert np.abs(got - CO.simulate_c(validate_circuit_dict(cd))).max() <= 1e-12


def test_jit_kernels_are_cached_by_structure():
    """Two circuits with the same structure but different angles share compiled kernels."""
    import ctypes as C
    from quantum_simulations_b200 import _lib as L
    from quantum_simulations_b200.kernel.cuda_dense import simulate

    def stats():
        v = [C.c_int() for _ in range(4)]
        s = C.c_double()
        L.load().qsv_jit_stats(*[C.byref(x) for x in v], C.byref(s))
        return [x.value for x in v]

    n = 16
    def circ(theta):
        gates = []
        for layer in range(4):
            gates += [{"qubits": [q], "gate": "RY", "params": {"theta": theta + 0.02 * q + 0.1 * layer}} for q in range(n)]
            gates += [{"qubits": [q, q + 1], "gate": "CZ"} for q in range(layer % 2, n - 1, 2)]
        return {"number_of_qubits": n, "gates": gates}
    a = simulate(circ(0.3))        # all angles stay in (0, pi/2): same ZYZ structure
    c0 = stats()
    b = simulate(circ(0.5))
    c1 = stats()
    assert c1[0] == c0[0] and c1[3] == 0, f"recompiled for new angles: {c0} -> {c1}"
    assert np.abs(a - CO.simulate_c(validate_circuit_dict(circ(0.3)))).max() <= 1e-12
    assert np.abs(b - CO.simulate_c(validate_circuit_dict(circ(0.5)))).max() <= 1e-12


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("n", [6, 12, 17])
def test_sample_indices_bit_exact(n, dtype):
    """qsv_sample reproduces oracle.sample_indices bit for bit on the same state and seed."""
    from quantum_simulations_b200.kernel.cuda import DeviceState
    rng = np.random.default_rng(n)
    psi = (rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)).astype(dtype)
    psi /= np.linalg.norm(psi)
    with DeviceState(n, dtype) as st:
        st.upload(psi)
        got = st.sample(seed=1234, shots=300)
    want = O.sample_indices(psi, 1234, 300)
    assert got.dtype == np.uint64 and np.array_equal(got, want)
    # a GHZ state samples only its two basis states
    from quantum_simulations_b200.kernel.cuda_dense import compile_circuit
    if dtype == "complex128" and n >= 12:
        with DeviceState(n) as st:
            st.init_zero(); st.run_program(compile_circuit(W.ghz(n)))
            s = st.sample(seed=5, shots=64)
        assert set(s.tolist()) <= {0, (1 << n) - 1} and len(set(s.tolist())) == 2


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_observables_match_oracle(dtype):
    """Marginal probabilities and <Z..Z> of a random state, incl. a sharded handle (rank bits)."""
    from quantum_simulations_b200.kernel.cuda import DeviceState
    n = 14
    rng = np.random.default_rng(3)
    psi = (rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)).astype(dtype)
    psi /= np.linalg.norm(psi)
    tol = 1e-12 if dtype == "complex128" else 1e-6
    with DeviceState(n, dtype) as st:
        st.upload(psi)
        for qs in ([0], [3, 9], [13, 0, 7, 2], list(range(12)), []):
            assert np.abs(st.probabilities(qs) - O.marginal_probabilities(psi, qs)).max() <= tol
        for qs in ([5], [0, 13], [1, 2, 3, 4, 10]):
            assert abs(st.expect_z(qs) - O.expectation_z(psi, qs)) <= tol
    world, n_local = 4, n - 2
    acc_p, acc_z = np.zeros(8), 0.0
    for rank in range(world):
        with DeviceState(n, dtype, rank=rank, world=world) as st:
            st.upload(psi[rank << n_local:(rank + 1) << n_local])
            acc_p += st.probabilities([13, 1, 12])
            acc_z += st.expect_z([12, 13, 4])
    assert np.abs(acc_p - O.marginal_probabilities(psi, [13, 1, 12])).max() <= tol
    assert abs(acc_z - O.expectation_z(psi, [12, 13, 4])) <= tol


def test_project_partial_measurement():
    """DeviceState.project == HiSVSIM-style collapse: keep one outcome of a qubit, renormalise."""
    from quantum_simulations_b200.kernel.cuda import DeviceState
    n = 12
    rng = np.random.default_rng(9)
    psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    psi /= np.linalg.norm(psi)
    for q, b in ((0, 1), (7, 0), (11, 1)):
        keep = ((np.arange(1 << n) >> q) & 1) == b
        p_want = float(np.sum(np.abs(psi[keep]) ** 2))
        want = np.where(keep, psi, 0) / np.sqrt(p_want)
        with DeviceState(n) as st:
            st.upload(psi)
            p = st.project(q, b)
            got = st.download()
            assert abs(st.norm2() - 1.0) < 1e-12
        assert abs(p - p_want) < 1e-12 and np.abs(got - want).max() < 1e-12
    with DeviceState(3) as st:
        st.init_zero()
        with pytest.raises(ValueError, match="probability 0"):
            st.project(1, 1)


def test_ghz20_config0_known_answer():
    from quantum_simulations_b200.kernel.cuda_dense import simulate
    got = simulate(W.ghz(20))
    s2 = 1 / np.sqrt(2)
    assert abs(got[0] - s2) < 1e-12 and abs(got[-1] - s2) < 1e-12
    assert np.count_nonzero(np.abs(got) > 1e-12) == 2


def test_program_replay_matches_one_shot():
    cuda = _cuda()
    from quantum_simulations_b200.kernel.cuda_dense import compile_circuit
    cd = W.random_1q_cz(16, 10, 5)
    prog = compile_circuit(cd)
    with cuda.DeviceState(16) as st:
        st.init_zero(); st.run_program(prog); a = st.download()
        h = st.upload_program(prog)
        st.init_zero(); st.replay(h); b = st.download()
        st.init_zero(); st.replay(h); c = st.download()
    assert np.array_equal(a, b) and np.array_equal(b, c)
    assert np.abs(a - CO.simulate_c(validate_circuit_dict(cd))).max() <= 1e-12


# ------------------------------------------------------- size-independent properties
@pytest.mark.parametrize("n", [24, 27, 30])
def test_large_invariants(n):
    """Sizes the NumPy oracle cannot reach cheaply: norm, QFT|0> uniform, GHZ, U U^dagger."""
    cuda = _cuda()
    from quantum_simulations_b200.kernel.cuda_dense import compile_circuit, circuit_ops
    from quantum_simulations_b200.circuit.passes import PassCompiler
    with cuda.DeviceState(n) as st:
        # QFT of |0> is the uniform superposition
        st.init_zero(); st.run_program(compile_circuit(W.qft(n)))
        assert abs(st.norm2() - 1.0) < 1e-10
        head = st.download(count=1 << 12)
        assert np.abs(head - 2.0 ** (-n / 2)).max() < 1e-12
        # GHZ
        st.init_zero(); st.run_program(compile_circuit(W.ghz(n)))
        assert abs(st.norm2() - 1.0) < 1e-12
        assert abs(st.download(count=1)[0] - 1 / np.sqrt(2)) < 1e-12
        assert abs(st.download(offset=(1 << n) - 1, count=1)[0] - 1 / np.sqrt(2)) < 1e-12
        # random circuit followed by its inverse returns |0...0>
        cd = validate_circuit_dict(W.random_1q_cz(n, 8, 77))
        ops = circuit_ops(cd)
        inv = [(qs, U.conj().T) for qs, U in reversed(ops)]
        st.init_zero(); st.run_program(PassCompiler(n).compile(ops + inv))
        assert abs(st.norm2() - 1.0) < 1e-10
        assert abs(st.download(count=1)[0] - 1.0) < 1e-10


def test_baseline_config_full_size_round_trip():
    """BASELINE.json configs[2] at FULL size (30 qubits, complex128, depth 20, seed 1234): the circuit
    followed by its inverse must return |0...0>, the forward state must be normalised, and the
    measurement samples of a GHZ-30 state are its two basis states."""
    cuda = _cuda()
    from quantum_simulations_b200.kernel.cuda_dense import circuit_ops, compile_circuit
    from quantum_simulations_b200.circuit.passes import PassCompiler
    n = 30
    cd = validate_circuit_dict(W.random_1q_cz(n, 20, 1234))
    ops = circuit_ops(cd)
    inv = [(qs, U.conj().T) for qs, U in reversed(ops)]
    with cuda.DeviceState(n) as st:
        st.init_zero(); st.run_program(PassCompiler(n).compile(ops))
        assert abs(st.norm2() - 1.0) < 1e-10
        st.run_program(PassCompiler(n).compile(inv))
        assert abs(st.norm2() - 1.0) < 1e-10
        assert abs(st.download(count=1)[0] - 1.0) < 1e-9
        st.init_zero(); st.run_program(compile_circuit(W.ghz(n)))
        s = st.sample(seed=3, shots=32)
        assert set(s.tolist()) == {0, (1 << n) - 1}


def test_n26_matches_c_oracle_sampled():
    from quantum_simulations_b200.kernel.cuda_dense import simulate
    n = 26
    cd = W.random_1q_cz(n, 20, 1234)
    want = CO.simulate_c(validate_circuit_dict(cd))
    got = simulate(cd)
    assert np.abs(got - want).max() <= 1e-12


def _full_vector_vs_c_oracle(n, dtype="complex128"):
    """Every amplitude of the BASELINE random depth-20 circuit at n qubits against oracle/ref_dense_c.c on the
    box's host cores (the reference's own kernels-vs-oracle test at its full size,
    wenbo_engine/tests/test_kernel_vs_ref.py:26-32; tolerance = BASELINE.json's)."""
    from quantum_simulations_b200.kernel.cuda_dense import simulate
    cd = W.random_1q_cz(n, 20, 1234)
    got = simulate(cd, dtype=dtype)
    want = CO.simulate_c(validate_circuit_dict(cd))
    worst = 0.0
    step = 1 << 24                                   # blockwise: no third full-size temporary on the host
    for o in range(0, 1 << n, step):
        worst = max(worst, float(np.abs(got[o:o + step] - want[o:o + step]).max()))
    print(f"full-vector parity n={n} {dtype}: max|d| = {worst:.3e} over 2^{n} amplitudes")
    assert worst <= TOL[dtype]


@pytest.mark.gpu
def test_n28_full_vector_matches_c_oracle():
    _full_vector_vs_c_oracle(28)


@pytest.mark.gpu
@pytest.mark.skipif(__import__("os").environ.get("QSV_TEST_FULL") != "1",
                    reason="BASELINE configs[2] at full size: 2 x 16 GiB on the host and ~2 min of the C oracle on 16 "
                           "cores — run with QSV_TEST_FULL=1 (log: profiles/r02/parity_n30_full_vector.log)")
def test_n30_full_vector_matches_c_oracle():
    _full_vector_vs_c_oracle(30)


# ----------------------------------------------------------------------- runner
@pytest.mark.parametrize("spec,chunk_size,kw", RUNNER_SPECS)
def test_runner_matches_reference_runner_golden(golden_runner, spec, chunk_size, kw):
    from quantum_simulations_b200.runner.single_node import run, collect_state
    key = f"runner/{spec}/cs{chunk_size}/" + ",".join(f"{k}={v}" for k, v in sorted(kw.items()))
    want = golden_runner[key]                        # reference stores complex64: atol 1e-6
    cd = circuit_from_spec(spec)
    with tempfile.TemporaryDirectory() as td:
        final = run(cd, td, chunk_size=chunk_size, **kw)
        got = collect_state(final, apply_permutation=True, work_dir=td)
    assert got.dtype == np.complex128
    assert np.abs(got - want).max() <= 1e-6
    assert np.abs(got - O.simulate(validate_circuit_dict(cd))).max() <= 1e-12


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
def test_runner_dtypes_and_manifest(dtype):
    from quantum_simulations_b200.runner.single_node import run, collect_state
    from quantum_simulations_b200.storage.manifest import read_manifest
    cd = W.qft(10)
    with tempfile.TemporaryDirectory() as td:
        final = run(cd, td, chunk_size=256, use_fusion=True, dtype=dtype)
        m = read_manifest(final)
        assert (m.dtype, m.n_chunks, m.chunk_size) == (dtype, 4, 256)
        got = collect_state(final)
    assert np.abs(got - O.simulate(validate_circuit_dict(cd))).max() <= TOL[dtype]


def test_pipeline_runner_surface(tmp_path):
    """reference tests/test_out_of_core_e2e.py:38 / test_fusion.py:149: pipeline.run == single_node.run."""
    from quantum_simulations_b200.runner import pipeline
    from quantum_simulations_b200.runner.single_node import collect_state
    cd = W.qft(6)
    for fusion in (False, True):
        final = pipeline.run(cd, tmp_path / f"f{int(fusion)}", chunk_size=16, buffer_depth=2, use_wal=False, use_fusion=fusion)
        assert np.abs(collect_state(final) - O.simulate(validate_circuit_dict(cd))).max() <= 1e-12


def test_qasm_front_end_on_the_device():
    """A QASMBench-style QFT (u1 / cx ladders) through simulate_qasm equals the reference QFT."""
    import math
    from quantum_simulations_b200.kernel.cuda_dense import simulate_qasm
    n = 16
    src = ['OPENQASM 2.0;', 'include "qelib1.inc";', f'qreg q[{n}];']
    for j in range(n):
        src.append(f"h q[{j}];")
        for k in range(j + 1, n):
            lam = 2 * math.pi / 2 ** (k - j + 1)
            src += [f"u1({lam / 2}) q[{k}];", f"cx q[{k}],q[{j}];", f"u1({-lam / 2}) q[{j}];", f"cx q[{k}],q[{j}];",
                    f"u1({lam / 2}) q[{j}];"]
    src += ["ccx q[0],q[1],q[2];", "rx(0.3) q[5];", "cswap q[3],q[4],q[5];"]
    got = simulate_qasm("\n".join(src))
    from quantum_simulations_b200.circuit.qasm import qasm_to_ops
    _, ops = qasm_to_ops("\n".join(src))
    want = np.zeros(1 << n, dtype=np.complex128)
    want[0] = 1
    O.apply_ops(want, ops)
    assert np.abs(got - want).max() <= 1e-12


def test_runner_errors():
    from quantum_simulations_b200.runner.single_node import run
    with tempfile.TemporaryDirectory() as td:
        with pytest.raises(ValueError, match="divisible by chunk_size"):
            run(W.qft(4), td, chunk_size=3)
        with pytest.raises(ValueError, match="kernel="):
            run(W.qft(4), td, kernel="scalar")
        with pytest.raises(ValueError, match="missing required keys"):
            run({"gates": []}, td)


@pytest.mark.parametrize("jit", [True, False])
@pytest.mark.parametrize("workload", ["random_1q_cz", "random_mixed", "qft"])
def test_low_position_register_stores(workload, jit):
    """PassCompiler(low_store_round=False): the last round may hold a content in registers that is stored
    to a low (128-byte row) position — no idle round before the store.  Same state, fewer rounds."""
    from quantum_simulations_b200.kernel.cuda_dense import compile_circuit, simulate
    n = 18
    cd = {"random_1q_cz": lambda: W.random_1q_cz(n, 20, 1234), "random_mixed": lambda: W.random_mixed(n, 400, 5),
          "qft": lambda: W.qft(n)}[workload]()
    assert compile_circuit(cd, low_store_round=False).stats["rounds"] <= compile_circuit(cd).stats["rounds"]
    want = CO.simulate_c(validate_circuit_dict(cd))
    for dtype in ("complex128", "complex64"):
        got = simulate(cd, dtype=dtype, jit=jit, low_store_round=False)
        assert np.abs(got - want).max() <= TOL[dtype]


def test_warp_local_rounds_on_the_device(monkeypatch):
    """warp_local_rounds + QSV_JIT_WARP_SYNC=1: __syncwarp() instead of the group barrier between rounds
    whose exchange stays inside each warp (tests/test_jit_host.py checks the same kernels on the CPU)."""
    from quantum_simulations_b200.kernel.cuda_dense import simulate
    monkeypatch.setenv("QSV_JIT_WARP_SYNC", "1")
    n = 20
    for cd in (W.random_1q_cz(n, 20, 1234), W.random_mixed(n, 400, 5), W.qft(n)):
        want = CO.simulate_c(validate_circuit_dict(cd))
        for dtype in ("complex128", "complex64"):
            got = simulate(cd, dtype=dtype, warp_local_rounds=True, low_store_round=False)
            assert np.abs(got - want).max() <= TOL[dtype]


def test_qasm_trajectories_with_measure_reset_if_on_the_device():
    """run_qasm (mid-circuit measure / reset / classical if, one seeded trajectory) against the oracle's
    definition: same outcomes, same classical registers, same final state."""
    import math
    from oracle import ref_dense as O
    from quantum_simulations_b200.circuit.qasm import qasm_to_steps
    from quantum_simulations_b200.kernel.cuda_dense import run_qasm
    head = 'OPENQASM 2.0;\ninclude "qelib1.inc";\n'
    teleport = head + ("qreg q[3]; creg c0[1]; creg c1[1]; ry(0.9) q[0]; t q[0]; h q[1]; cx q[1],q[2]; cx q[0],q[1]; h q[0];"
                       "measure q[0] -> c0[0]; measure q[1] -> c1[0]; if(c1==1) x q[2]; if(c0==1) z q[2]; reset q[0]; reset q[1];")
    n = 14
    big = head + f"qreg q[{n}]; creg c[3];" + "".join(f"h q[{i}];" for i in range(n)) + \
        "".join(f"cx q[{i}],q[{i + 1}]; rz({0.1 * (i + 1)}) q[{i + 1}];" for i in range(n - 1)) + \
        "measure q[3] -> c[0]; measure q[9] -> c[1]; if(c==3) ry(0.4) q[0]; if(c==1) h q[5]; reset q[9];" + \
        "".join(f"ry({0.05 * (i + 1)}) q[{i}];" for i in range(n)) + "measure q[0] -> c[2]; if(c==5) x q[13];"
    for text in (teleport, big):
        nq, steps, cregs = qasm_to_steps(text)
        for seed in range(4):
            want, wbits = O.run_qasm_steps(nq, steps, cregs, seed)
            got, gbits = run_qasm(text, seed=seed)
            assert gbits == wbits
            assert np.abs(got - want).max() <= 1e-12
    a, b = math.cos(0.45), math.sin(0.45) * np.exp(0.25j * math.pi)
    got, _ = run_qasm(teleport, seed=11)
    want = np.zeros(8, dtype=np.complex128)
    want[0], want[4] = a, b
    assert abs(abs(np.vdot(want, got)) - 1) < 1e-12
