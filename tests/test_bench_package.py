"""quantum_simulations_b200.bench: the MQT-Bench runner mirror (wenbo_engine/bench/mqt_bench_runner.py), the per-gate
kernel table (bench/kernel.py) and the end-to-end table (bench/end_to_end.py).

CPU: the native families against the oracle and against their closed forms, the planner on them through the pass
emulator, the reference's benchmark table, the control flow of the kernel table with a stand-in device.
GPU: the runner's own correctness column (closed form + fused path vs per-gate path), the tables on a small state."""
import io

import numpy as np
import pytest

from oracle import ref_dense as O
from quantum_simulations_b200.bench import kernel as KB
from quantum_simulations_b200.bench import mqt_bench_runner as MQ
from quantum_simulations_b200.circuit.io import validate_circuit_dict
from quantum_simulations_b200.circuit.sharding import plan_single
from quantum_simulations_b200.kernel.cuda_dense import circuit_ops
from tests.pass_emulator import run_program

CLOSED = ["ghz", "graphstate", "wstate", "bv", "dj", "qft", "qftentangled", "qpeexact"]


@pytest.mark.parametrize("family", CLOSED)
@pytest.mark.parametrize("n", [3, 4, 5, 6, 8, 11])
def test_closed_forms_are_what_the_oracle_computes(family, n):
    cd = validate_circuit_dict(MQ.native_circuit(family, n))
    want = MQ.expected_state(family, n)
    got = O.simulate(cd)
    assert want is not None and abs(np.vdot(want, want) - 1) < 1e-12
    assert np.abs(got - want).max() <= 1e-12


@pytest.mark.parametrize("family", MQ.NATIVE_FAMILIES)
@pytest.mark.parametrize("n", [4, 7, 10])
def test_native_families_plan_into_passes(family, n):
    """every family through the planner the runner uses, executed by the NumPy pass emulator"""
    cd = validate_circuit_dict(MQ.native_circuit(family, n))
    prog = plan_single(circuit_ops(cd), n, "complex128", True, False, tile_bits=min(n, 6), low_bits=2)
    psi = np.zeros(1 << n, dtype=np.complex128)
    if not prog.fused_init:
        psi[0] = 1
    run_program(prog, psi)
    assert prog.final_pos == list(range(n))
    assert np.abs(psi - O.simulate(cd)).max() <= 1e-12


def test_native_generators_are_seeded_and_sized():
    a, b = MQ.native_circuit("randomcircuit", 9), MQ.native_circuit("randomcircuit", 9)
    assert a == b and a != MQ.native_circuit("randomcircuit", 9, seed=11)
    assert len(MQ.native_circuit("ghz", 20)["gates"]) == 20
    assert len(MQ.native_circuit("qft", 16)["gates"]) == 16 + 16 * 15 // 2
    with pytest.raises(KeyError):
        MQ.native_circuit("shor", 18)
    assert MQ.expected_state("realamp", 5) is None


def test_benchmark_table_is_the_references():
    """wenbo_engine/bench/mqt_bench_runner.py:28-68: 31 families, the same sizes"""
    t = dict(MQ.BENCHMARKS)
    assert len(MQ.BENCHMARKS) == 31 and len(t) == 31
    assert t["ghz"] == [3, 4, 5, 6, 8, 10, 14, 18, 20] and t["shor"] == [18] and t["qpeexact"] == [3, 4, 5, 6, 8, 10, 12, 14]
    assert t["randomcircuit"] == [3, 4, 5, 6, 8, 10, 12, 14] and t["hrs_cumulative_multiplier"] == [5]
    assert MQ.CORRECTNESS_MAX_N == 20
    ref = None
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location("ref_mqt", "/root/reference/wenbo_engine/bench/mqt_bench_runner.py")
        src = open(spec.origin).read()
        ns: dict = {}
        exec(src[src.index("BENCHMARKS = ["):src.index("CORRECTNESS_MAX_N")], ns)
        ref = ns["BENCHMARKS"]
    except (OSError, ValueError, AttributeError):
        pass                                               # the reference tree is not on the GPU box
    if ref is not None:
        assert ref == MQ.BENCHMARKS


def test_without_qiskit_the_runner_says_so_and_runs_native(monkeypatch, capsys):
    assert MQ.HAS_DEPS is False                             # qiskit / mqt.bench are not in this image
    seen = {}
    monkeypatch.setattr(MQ, "run_native", lambda rows, fam, max_n, perf: seen.update(fam=fam, max_n=max_n, perf=perf) or 0)
    assert MQ.main(["--max-n", "12", "--families", "ghz", "--no-perf"]) == 0
    assert seen == {"fam": ["ghz"], "max_n": 12, "perf": False}
    assert "not installed" in capsys.readouterr().out


class _FakeState:
    """stands in for DeviceState: counts calls, 'times' every region as 1 ms per call"""
    def __init__(self, n, dtype="complex128", device=0):
        self.n_qubits, self.dtype, self.n_amps, self.calls, self._k = n, np.dtype(dtype), 1 << n, [], 0
    def __enter__(self): return self
    def __exit__(self, *a): return False
    def init_zero(self): pass
    def sync(self): pass
    def timer_start(self): self._k = len(self.calls)
    def timer_stop(self): return float(len(self.calls) - self._k)
    def __getattr__(self, name):
        if name.startswith(("apply_", "norm2", "probabilities", "expect_z", "sample")):
            return lambda *a, **k: self.calls.append(name) or 1.0
        raise AttributeError(name)


def test_kernel_table_control_flow_and_accounting(monkeypatch):
    monkeypatch.setattr(KB, "DeviceState", _FakeState)
    out = io.StringIO()
    rows = KB.bench_kernel(12, "complex128", reps=3, out=out)
    kinds = {r["kernel"] for r in rows}
    assert kinds == {"k_apply_1q", "k_apply_2q", "k_apply_ctrl_1q", "k_apply_diag", "k_apply_kq", "k_norm2_partial",
                     "probabilities", "expect_z", "sample"}
    peak, _ = KB.hbm_peak_gbs()
    for r in rows:
        assert r["ms"] == 1.0                               # 3 calls / 3 reps
        touched = 0.5 if r["kernel"] == "k_apply_ctrl_1q" else 1.0
        writes = r["kernel"].startswith("k_apply")
        assert r["algorithmic_bytes"] == int((2 if writes else 1) * 16 * touched * 4096)
        assert r["frac"] == pytest.approx(r["gbs"] / peak, abs=1e-3)
    assert "H on qubit 0" in out.getvalue() and "norm^2" in out.getvalue()
    quick = KB.bench_kernel(12, "complex128", reps=1, out=io.StringIO(), quick=True)      # what bench.py embeds in its line
    assert {r["kernel"] for r in quick} == kinds and len(quick) < len(rows) // 2
    rows64 = KB.bench_kernel(12, "complex64", reps=1, observables=False, out=io.StringIO())
    assert all(r["dtype"] == "complex64" and not r["kernel"].startswith(("prob", "expect", "sample")) for r in rows64)


def test_io_table_runs_on_the_chunk_store():
    from quantum_simulations_b200.bench import io as BIO
    r = BIO.bench_io(1 << 12, 3, out=io.StringIO())
    assert r["write_MBs"] > 0 and r["read_MBs"] > 0


def test_hyperparam_sweep_table(monkeypatch):
    """the reference's sweep (bench/hyperparam_sweep.py:33-118): every (chunk, fusion) for the single-node runner and
    every buffer depth for the pipeline runner, without WAL"""
    from quantum_simulations_b200.bench import hyperparam_sweep as HS
    from quantum_simulations_b200 import workloads as W
    calls = []
    monkeypatch.setattr(HS, "sn_run", lambda cd, td, **kw: calls.append(("sn", kw)))
    monkeypatch.setattr(HS, "pl_run", lambda cd, td, **kw: calls.append(("pl", kw)))
    out = io.StringIO()
    res = HS.sweep(lambda: W.qft(6), "QFT-6", chunk_exponents=[4, 8], buffer_depths=[1, 4], reps=2, out=out)
    assert len(res) == 2 * 2 * (1 + 2) and len(calls) == 2 * len(res)
    assert all(kw["use_wal"] is False for _, kw in calls)
    assert {kw["chunk_size"] for _, kw in calls} == {16, 64}            # 2^8 is clipped to the state (2^6)
    assert {kw.get("buffer_depth") for k, kw in calls if k == "pl"} == {1, 4}
    assert "HYPERPARAMETER SWEEP: QFT-6" in out.getvalue() and "best:" in out.getvalue()


def test_matmul_vs_io_table():
    """the reference's side-by-side table (bench/matmul_vs_io.py:84-141) with the device tier stubbed: files tier real"""
    from quantum_simulations_b200.bench import matmul_vs_io as MV
    seen = []

    def fake_tiers(cs, dtype, device):
        seen.append(cs)
        return {"1q_GBs": 4000.0, "2q_GBs": 6000.0, "ms_per_gate": 0.01, "pcie_GBs": 50.0, "pcie_ms_per_chunk": 1.0}

    out = io.StringIO()
    rows = MV.bench_compare([1 << 10, 1 << 12], out=out, tiers=fake_tiers)
    assert seen == [1 << 10, 1 << 12] and len(rows) == 2
    assert all(r["kernel_over_pcie"] == 100.0 and r["gates_to_match_pcie"] == 100 and r["files_MBs"] > 0 for r in rows)
    assert "I/O bound (fuse!)" in out.getvalue() and "gates to match" in out.getvalue()
    assert [MV.verdict(x) for x in (11, 5, 1)] == ["I/O bound (fuse!)", "I/O leaning", "balanced"]
    with pytest.raises(ValueError):
        MV.device_tiers(3000)


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("family", MQ.NATIVE_FAMILIES)
def test_runner_correctness_column_on_the_device(family):
    for n in (3, 6, 12):
        cd = MQ.native_circuit(family, n)
        ok, o_exact, o_paths = MQ._check_native(family, cd)
        assert ok, (family, n, o_exact, o_paths)
        assert abs(o_paths - 1) < 1e-9 and (o_exact is None or abs(o_exact - 1) < 1e-9)
        # and the oracle agrees with what the runner accepted
        assert np.abs(MQ._simulate(cd) - O.simulate(validate_circuit_dict(cd))).max() <= 1e-12


@pytest.mark.gpu
def test_runner_tables_on_the_device(capsys):
    rows: list = []
    assert MQ.run_native(rows, ["ghz", "qpeexact"], max_n=10, perf=True) == 0
    assert [r["correct"] for r in rows] == ["PASS"] * len(rows) and all(r["time_s"] > 0 and r["device_s"] > 0 for r in rows)
    krows = KB.bench_kernel(20, "complex128", reps=2, out=io.StringIO())
    assert all(r["ms"] > 0 and r["gbs"] > 0 for r in krows)
    from quantum_simulations_b200.bench import end_to_end as E
    from quantum_simulations_b200 import workloads as W
    r = E.bench_e2e(lambda: W.qft(10), "QFT-10")
    assert set(r) == {"single_node", "pipeline"} and all(v["time"] > 0 for v in r.values())
