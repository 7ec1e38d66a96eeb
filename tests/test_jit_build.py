"""The run-time specialisation path builds without a GPU: generate + NVRTC (sm_100a) for the
passes of real compiled programs.  No kernel is launched here."""
import ctypes as C

import pytest

from quantum_simulations_b200 import _lib as L, workloads as W
from quantum_simulations_b200.kernel.cuda_dense import compile_circuit


def _build(step, dtype=L.QSV_C128):
    n, log = C.c_size_t(), C.create_string_buffer(1 << 14)
    rc = L.load().qsv_jit_build_pass(C.byref(step.desc), step.ops, dtype, C.byref(n), log, len(log))
    return rc, n.value, log.value.decode(errors="replace")


@pytest.mark.parametrize("circuit", ["random_1q_cz", "qft", "random_mixed"])
def test_every_ring_pass_specialises(circuit):
    cd = {"random_1q_cz": lambda: W.random_1q_cz(16, 20, 1234), "qft": lambda: W.qft(14),
          "random_mixed": lambda: W.random_mixed(14, 200, 3)}[circuit]()
    prog = compile_circuit(cd)
    assert prog.passes
    for step in prog.passes[:3]:
        rc, size, log = _build(step)
        if rc == L.QSV_EIO:
            pytest.skip(f"NVRTC not installed: {log}")
        assert rc == 0 and size > 10000, log


def test_complex64_passes_specialise_too():
    prog = compile_circuit(W.random_1q_cz(16, 20, 1234), dtype="complex64")
    assert prog.stats["tile_bits"] == 11
    for step in prog.passes[:2]:
        rc, size, log = _build(step, L.QSV_C64)
        if rc == L.QSV_EIO:
            pytest.skip(f"NVRTC not installed: {log}")
        assert rc == 0 and size > 10000, log


def test_other_tile_sizes_are_left_to_the_interpreter():
    prog = compile_circuit(W.random_1q_cz(12, 6, 1), tile_bits=8)
    rc, size, log = _build(prog.passes[0])
    assert rc == L.QSV_EINVAL and "not eligible" in log
