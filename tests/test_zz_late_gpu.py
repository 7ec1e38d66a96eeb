"""GPU tests written after the round's last GPU minute.  They sort LAST on purpose: the driver runs `pytest -m gpu -x`, and a
surprise in a test that has never met the hardware must not hide the validated suite behind it."""
import io

import pytest


@pytest.mark.gpu
def test_sweep_and_pinned_copy_on_the_device():
    from quantum_simulations_b200.bench import hyperparam_sweep as HS, io as BIO
    from quantum_simulations_b200 import workloads as W
    res = HS.sweep(lambda: W.qft(10), "QFT-10", chunk_exponents=[8], buffer_depths=[1, 2], out=io.StringIO())
    assert len(res) == 6 and all(r[4] > 0 for r in res)
    assert BIO.bench_pinned(16, out=io.StringIO())["d2h_GBs"] > 0


@pytest.mark.gpu
def test_deep_circuit_runs_on_specialised_kernels_only():
    """PassCompiler(fit_jit=True): a rotation-heavy deep circuit whose greedy passes exceed the specialised kernel's limits
    is re-planned into passes that fit; on the device every pass is specialised (none interpreted) and the state equals
    the oracle's."""
    import numpy as np
    from oracle import ref_dense as O
    from quantum_simulations_b200.circuit import sharding
    from quantum_simulations_b200.circuit.passes import fits_jit
    from quantum_simulations_b200.kernel.cuda_dense import circuit_ops, simulate
    from tests.test_fit_jit import _deep_rotations
    n = 13
    cd = _deep_rotations(n, 14, 2)
    prog = sharding.plan_single(circuit_ops(cd), n, "complex128", True, False)
    assert all(fits_jit(s) for s in prog.passes) and prog.stats.get("fit_jit_max_ops")
    got = simulate(cd)
    assert np.abs(got - O.simulate(cd)).max() <= 1e-11


@pytest.mark.gpu
def test_matmul_vs_io_on_the_device():
    from quantum_simulations_b200.bench import matmul_vs_io as MV
    rows = MV.bench_compare([1 << 16, 1 << 20], out=io.StringIO())
    assert len(rows) == 2 and all(r["1q_GBs"] > 0 and r["2q_GBs"] > 0 and r["pcie_GBs"] > 0 and r["files_MBs"] > 0 for r in rows)
