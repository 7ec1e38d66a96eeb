"""Drop-in GPU forms of the reference's butterfly kernels for gates on "non-local" qubits
(wenbo_engine/kernel/cpu_nonlocal.py:22-67): same names, same arguments, in place on the
host chunks the caller loaded.

On a GPU a qubit that selects the CHUNK is just one more index bit: the partner chunks are
uploaded side by side into one device state of log2(len) + 1 (pair) or + 2 (quad) qubits and
the gate runs as an ordinary ``qsv_apply_1q / qsv_apply_2q`` on that top bit.  These wrappers
exist for the reference's chunk-at-a-time runner (one PCIe round trip per call); the resident
path (runner.single_node / runner.multi_gpu) never needs them."""
from __future__ import annotations

import math

import numpy as np

from quantum_simulations_b200.kernel.cuda import DeviceState


def _check(chunks):
    n0 = len(chunks[0])
    k = int(math.log2(n0))
    if 1 << k != n0 or any(len(c) != n0 or c.dtype != chunks[0].dtype or c.ndim != 1 for c in chunks):
        raise ValueError("partner chunks must be 1-D arrays of the same power-of-two length and dtype")
    if chunks[0].dtype not in (np.complex64, np.complex128):
        raise ValueError("chunks must be complex64 or complex128")
    return k


def _round_trip(chunks, fn) -> None:
    k = _check(chunks)
    extra = int(math.log2(len(chunks)))
    with DeviceState(k + extra, chunks[0].dtype) as st:
        for i, c in enumerate(chunks):
            st.upload(c, offset=i << k)
        fn(st, k)
        for i, c in enumerate(chunks):
            if c.flags.c_contiguous:
                st.download(c, offset=i << k, count=1 << k)
            else:
                c[:] = st.download(offset=i << k, count=1 << k)


def apply_1q_pair(c0: np.ndarray, c1: np.ndarray, U: np.ndarray) -> None:
    """1-qubit gate across two partner chunks (cpu_nonlocal.py:22-26)."""
    _round_trip([c0, c1], lambda st, k: st.apply_1q(k, U))


def apply_2q_pair_qa_local(c0: np.ndarray, c1: np.ndarray, qa: int, U: np.ndarray) -> None:
    """2-qubit gate, qa local, qb selects the chunk (cpu_nonlocal.py:29-42)."""
    _round_trip([c0, c1], lambda st, k: st.apply_2q(qa, k, U))


def apply_2q_pair_qb_local(c0: np.ndarray, c1: np.ndarray, qb: int, U: np.ndarray) -> None:
    """2-qubit gate, qa selects the chunk, qb local (cpu_nonlocal.py:45-58)."""
    _round_trip([c0, c1], lambda st, k: st.apply_2q(k, qb, U))


def apply_2q_quad(c00: np.ndarray, c01: np.ndarray, c10: np.ndarray, c11: np.ndarray, U: np.ndarray) -> None:
    """2-qubit gate, both qubits select the chunk: c01 has qb set, c10 has qa set
    (cpu_nonlocal.py:61-67, runner/single_node.py:315-320)."""
    # device layout: chunk index bit 0 = qb, bit 1 = qa  ->  order c00, c01, c10, c11
    _round_trip([c00, c01, c10, c11], lambda st, k: st.apply_2q(k + 1, k, U))
