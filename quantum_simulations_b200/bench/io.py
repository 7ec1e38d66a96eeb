"""Sequential I/O bandwidth of the chunk store (the checkpoint files of the runner) and of the pinned device-to-host
copies that feed it.

``bench_io`` is the reference's ``wenbo_engine/bench/io.py:13-40`` over this package's ``storage.block_store``
(same chunk files, same atomic write): MB/s of writing and reading ``n_chunks`` chunks.  ``bench_pinned`` is the half
of the checkpoint path the reference does not have: GB/s of ``qsv_download`` into pinned host memory (CUDA device
required).  On the GPU path chunk files are written only at checkpoints and at the end of a run, by a writer thread
beside the kernels (runner/pipeline.py), so these numbers bound the checkpoint interval, not the gate throughput.

    python -m quantum_simulations_b200.bench.io [--pinned-qubits 28]
"""
from __future__ import annotations

import argparse
import tempfile
import time
from pathlib import Path

import numpy as np

from quantum_simulations_b200.storage.block_store import DTYPE, read_chunk, write_chunk_atomic


def bench_io(chunk_size: int = 1 << 20, n_chunks: int = 16, out=None) -> dict:
    """Measure sequential read/write throughput of chunk files in MB/s."""
    bytes_per_chunk = chunk_size * np.dtype(DTYPE).itemsize
    total_bytes = bytes_per_chunk * n_chunks
    with tempfile.TemporaryDirectory() as td:
        p = Path(td)
        data = np.random.default_rng(0).standard_normal(2 * chunk_size).astype(np.float32).view(DTYPE)
        t0 = time.perf_counter()
        for i in range(n_chunks):
            write_chunk_atomic(p / f"chunk_{i:06d}.bin", data)
        t_write = time.perf_counter() - t0
        t0 = time.perf_counter()
        for i in range(n_chunks):
            got = read_chunk(p / f"chunk_{i:06d}.bin")
        t_read = time.perf_counter() - t0
        assert got.shape == data.shape and np.array_equal(got, data)
    mb = total_bytes / 1e6
    print(f"chunk_size={chunk_size}  n_chunks={n_chunks}  total={mb:.1f} MB", file=out)
    print(f"  write: {t_write:.3f}s  -> {mb / t_write:.1f} MB/s", file=out)
    print(f"  read:  {t_read:.3f}s  -> {mb / t_read:.1f} MB/s", file=out)
    return {"write_MBs": mb / t_write, "read_MBs": mb / t_read}


def bench_pinned(n_qubits: int = 28, dtype: str = "complex128", device: int = 0, reps: int = 3, out=None) -> dict:
    """GB/s of the device-to-host copy of a whole state into pinned memory (what a checkpoint moves first)."""
    from quantum_simulations_b200.kernel.cuda import DeviceState
    from quantum_simulations_b200.storage.pinned import PinnedBuffer
    nbytes = (1 << n_qubits) * np.dtype(dtype).itemsize
    host = PinnedBuffer(nbytes)
    try:
        arr = host.array(dtype, 1 << n_qubits)
        with DeviceState(n_qubits, dtype, device) as st:
            st.init_zero()
            st.download(arr)
            st.sync()
            t0 = time.perf_counter()
            for _ in range(reps):
                st.download(arr)
            dt = (time.perf_counter() - t0) / reps
    finally:
        host.free()
    print(f"pinned D2H of 2^{n_qubits} {dtype} ({nbytes / 2**30:.2f} GiB): {dt * 1e3:.1f} ms -> {nbytes / dt / 1e9:.1f} GB/s", file=out)
    return {"d2h_GBs": nbytes / dt / 1e9, "ms": dt * 1e3}


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--pinned-qubits", type=int, default=0, help="also time the pinned device-to-host copy of a 2^n state")
    a = ap.parse_args(argv)
    for cs_exp in (18, 20, 22):
        bench_io(chunk_size=1 << cs_exp, n_chunks=16)
        print()
    if a.pinned_qubits:
        bench_pinned(a.pinned_qubits)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
