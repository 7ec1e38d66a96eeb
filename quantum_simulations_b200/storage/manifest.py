"""manifest.json of one committed state buffer (reference storage/manifest.py:18-64,
docs/storage_spec.md:36-52).  Same fields; ``dtype`` gains "complex128" because the
GPU engine carries complex128 end to end — a complex64 manifest
written here is byte-compatible with the reference's reader."""
from __future__ import annotations

import json
import time
from dataclasses import asdict, dataclass, field
from pathlib import Path
from typing import List

from quantum_simulations_b200.storage._atomic import publish_text

SUPPORTED_DTYPES = ("complex64", "complex128")


@dataclass
class Manifest:
    n_qubits: int
    chunk_size: int          # amplitudes per chunk
    n_chunks: int
    dtype: str = "complex64"
    chunks: List[str] = field(default_factory=list)
    created: float = field(default_factory=time.time)

    def validate(self) -> None:
        if self.chunk_size * self.n_chunks != 1 << self.n_qubits:
            raise ValueError(f"chunk_size*n_chunks={self.chunk_size * self.n_chunks} "
                             f"!= 2^n_qubits={1 << self.n_qubits}")
        if len(self.chunks) != self.n_chunks:
            raise ValueError(f"chunk list length {len(self.chunks)} != n_chunks {self.n_chunks}")
        if self.dtype not in SUPPORTED_DTYPES:
            raise ValueError(f"unsupported dtype {self.dtype}")


def write_manifest_atomic(directory: str | Path, manifest: Manifest) -> Path:
    manifest.validate()
    return publish_text(Path(directory) / "manifest.json", json.dumps(asdict(manifest), indent=2))


def read_manifest(directory: str | Path) -> Manifest:
    m = Manifest(**json.loads((Path(directory) / "manifest.json").read_text()))
    m.validate()
    return m
