"""Gate kernels against the data paths that feed them, side by side, per chunk size.

GPU counterpart of the reference's ``wenbo_engine/bench/matmul_vs_io.py:22-141``.  There the question is how many
gates a chunk must receive per read + write of its file for the CPU to stay busy ("I/O bound: fuse!"); the answer is
what ``batch_levels`` exists for.  On the device the same question has two tiers:

* a gate kernel streams a chunk through HBM (``qsv_apply_1q`` / ``qsv_apply_2q``, one read + one write per gate);
* the chunk reaches the device over PCIe from pinned host memory (``qsv_upload`` / ``qsv_download``), and the host
  memory is filled from / drained to chunk files (``storage.block_store``, the reference's atomic chunk files).

The table gives, per chunk size, the throughput of each tier in the reference's accounting (chunk bytes per gate,
chunk bytes per transfer), the ratio kernel / PCIe and kernel / files, and "gates to match": how many gates a chunk must
receive per round trip over that tier for the kernels to be the longer side — the number the pass compiler's fusion
(circuit/passes.py: a pass applies tens of gates per sweep) and the checkpoint interval of the runners are chosen
against.

    python -m quantum_simulations_b200.bench.matmul_vs_io [--exponents 20 22 24 26] [--dtype complex64]
"""
from __future__ import annotations

import argparse
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

from quantum_simulations_b200.kernel import gates as gmod
from quantum_simulations_b200.storage.block_store import DTYPE, read_chunk, write_chunk_atomic


def files_rw(chunk_size: int, n_chunks: int = 4) -> dict:
    """write + read of `n_chunks` chunk files of `chunk_size` amplitudes (the store's own dtype): MB/s and ms per chunk"""
    data = np.random.default_rng(1).standard_normal(2 * chunk_size).astype(np.float32).view(DTYPE)
    mb = data.nbytes * n_chunks / 1e6
    with tempfile.TemporaryDirectory() as td:
        names = [Path(td) / f"c{i:06d}.bin" for i in range(n_chunks)]
        t0 = time.perf_counter()
        for f in names:
            write_chunk_atomic(f, data)
        t_w = time.perf_counter() - t0
        t0 = time.perf_counter()
        for f in names:
            got = read_chunk(f)
        t_r = time.perf_counter() - t0
        assert np.array_equal(got, data)
    return {"write_MBs": mb / t_w, "read_MBs": mb / t_r, "rw_MBs": 2 * mb / (t_w + t_r), "ms_per_chunk": (t_w + t_r) / n_chunks * 1e3}


def device_tiers(chunk_size: int, dtype: str = "complex128", device: int = 0, n_gates: int = 10, reps: int = 3) -> dict:
    """One device state of `chunk_size` amplitudes: H on qubit 0 and CNOT(0, 1) timed with CUDA events (the reference's
    two cases, bench/matmul_vs_io.py:54-76), and the pinned upload + download of the chunk timed on the host."""
    from quantum_simulations_b200.kernel.cuda import DeviceState
    from quantum_simulations_b200.storage.pinned import PinnedBuffer
    n = chunk_size.bit_length() - 1
    if 1 << n != chunk_size or n < 4:
        raise ValueError("chunk_size must be a power of two >= 16")
    nbytes = chunk_size * np.dtype(dtype).itemsize
    h, cx = gmod.gate_matrix("H", {}), gmod.gate_matrix("CNOT", {})
    host = PinnedBuffer(nbytes)
    try:
        arr = host.array(dtype, chunk_size)
        arr[:] = 0
        arr[0] = 1
        with DeviceState(n, dtype, device) as st:
            st.upload(arr)
            st.apply_1q(0, h)                                # warm-up of both kernels
            st.apply_2q(0, 1, cx)
            st.sync()

            def timed(fn) -> float:
                st.timer_start()
                for _ in range(n_gates * reps):
                    fn()
                return st.timer_stop() / (n_gates * reps)

            ms_1q = timed(lambda: st.apply_1q(0, h))
            ms_2q = timed(lambda: st.apply_2q(0, 1, cx))
            st.sync()
            t0 = time.perf_counter()
            for _ in range(reps):
                st.upload(arr)
                st.download(arr)
            ms_pcie = (time.perf_counter() - t0) / reps * 1e3
    finally:
        host.free()
    return {"1q_GBs": nbytes / ms_1q / 1e6, "2q_GBs": nbytes / ms_2q / 1e6, "ms_per_gate": (ms_1q + ms_2q) / 2,
            "pcie_GBs": 2 * nbytes / ms_pcie / 1e6, "pcie_ms_per_chunk": ms_pcie}


def verdict(ratio: float) -> str:
    """the reference's three classes (bench/matmul_vs_io.py:107-112)"""
    return "I/O bound (fuse!)" if ratio > 10 else "I/O leaning" if ratio > 2 else "balanced"


def bench_compare(chunk_sizes=None, dtype: str = "complex128", device: int = 0, out=sys.stdout, tiers=device_tiers) -> list[dict]:
    if chunk_sizes is None:
        chunk_sizes = [1 << e for e in (20, 22, 24, 26)]
    rows = []
    for cs in chunk_sizes:
        io, dv = files_rw(cs), tiers(cs, dtype, device)
        kern = (dv["1q_GBs"] + dv["2q_GBs"]) / 2
        rows.append({"chunk_size": cs, "chunk_MB": cs * np.dtype(dtype).itemsize / 1e6, "files_MBs": io["rw_MBs"],
                     "pcie_GBs": dv["pcie_GBs"], "1q_GBs": dv["1q_GBs"], "2q_GBs": dv["2q_GBs"],
                     "kernel_over_pcie": kern / max(dv["pcie_GBs"], 1e-12), "kernel_over_files": kern * 1e3 / max(io["rw_MBs"], 1e-12),
                     "gates_to_match_pcie": int(dv["pcie_ms_per_chunk"] / max(dv["ms_per_gate"], 1e-12)),
                     "gates_to_match_files": int(io["ms_per_chunk"] * np.dtype(dtype).itemsize / np.dtype(DTYPE).itemsize
                                                 / max(dv["ms_per_gate"], 1e-12))})
    print("GATE KERNELS vs THE DATA PATHS THAT FEED THEM", file=out)
    print(f"{'chunk_size':>12} {'chunk_MB':>9} {'files MB/s':>11} {'PCIe GB/s':>10} {'1Q GB/s':>9} {'2Q GB/s':>9} "
          f"{'k/PCIe':>8} {'k/files':>9} {'verdict (PCIe)':>18}", file=out)
    for r in rows:
        print(f"{r['chunk_size']:>12,} {r['chunk_MB']:>9.2f} {r['files_MBs']:>11.1f} {r['pcie_GBs']:>10.1f} {r['1q_GBs']:>9.1f} "
              f"{r['2q_GBs']:>9.1f} {r['kernel_over_pcie']:>7.1f}x {r['kernel_over_files']:>8.0f}x {verdict(r['kernel_over_pcie']):>18}", file=out)
    print(file=out)
    print(f"{'chunk_size':>12} {'gates to match PCIe':>20} {'gates to match files':>21}", file=out)
    for r in rows:
        print(f"{r['chunk_size']:>12,} {r['gates_to_match_pcie']:>20} {r['gates_to_match_files']:>21}", file=out)
    print("gates to match = gates a chunk must receive per round trip over that tier for the kernels to be the longer side;\n"
          "a fused pass applies tens of gates per sweep, and the runners touch the tiers only at checkpoints and at the end.", file=out)
    return rows


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--exponents", type=int, nargs="*", default=[20, 22, 24, 26])
    ap.add_argument("--dtype", default="complex128", choices=["complex64", "complex128"])
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args(argv)
    bench_compare([1 << e for e in a.exponents], a.dtype, a.device)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
