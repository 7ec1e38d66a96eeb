"""Chunk files of a state buffer: raw little-endian complex arrays ``chunk_%06d.bin``
(reference storage/block_store.py:11-65).  The reference fixes complex64; here the
dtype is a parameter (default complex64 keeps files interchangeable with it)."""
from __future__ import annotations

from pathlib import Path

import numpy as np

from quantum_simulations_b200.storage._atomic import publish_bytes
from quantum_simulations_b200.storage.manifest import Manifest, write_manifest_atomic

DTYPE = np.complex64


def chunk_filename(idx: int) -> str:
    return f"chunk_{idx:06d}.bin"


def write_chunk_atomic(path: str | Path, data: np.ndarray, dtype=DTYPE) -> None:
    arr = np.ascontiguousarray(data, dtype=dtype)
    publish_bytes(path, arr.view(np.uint8).data)


def read_chunk(path: str | Path, dtype=DTYPE) -> np.ndarray:
    return np.fromfile(str(path), dtype=dtype)


def init_zero_state(directory: str | Path, n_qubits: int, chunk_size: int = 1 << 20,
                    dtype=DTYPE) -> Manifest:
    """|0...0> on disk + manifest (reference block_store.py:35-65)."""
    total = 1 << n_qubits
    if total % chunk_size:
        raise ValueError("2^n_qubits must be divisible by chunk_size")
    root = Path(directory)
    names = [chunk_filename(c) for c in range(total // chunk_size)]
    for c, name in enumerate(names):
        block = np.zeros(chunk_size, dtype=dtype)
        if c == 0:
            block[0] = 1.0
        write_chunk_atomic(root / "chunks" / name, block, dtype)
    man = Manifest(n_qubits=n_qubits, chunk_size=chunk_size, n_chunks=len(names),
                   dtype=np.dtype(dtype).name, chunks=names)
    write_manifest_atomic(root, man)
    return man
