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
