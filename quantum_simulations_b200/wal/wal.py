"""Step-level write-ahead log, same file schema as the reference (wal/wal.py:17-93):

    wal.json = {"circuit_hash": sha256(normalised circuit)[:16],
                "committed_buf": "a" | "b", "done_steps": int}

rewritten atomically after every committed checkpoint.  In the GPU engine a "step" that
reaches disk is a checkpoint of the device shards, not every pass."""
from __future__ import annotations

import hashlib
import json
from pathlib import Path

from quantum_simulations_b200.circuit.io import validate_circuit_dict
from quantum_simulations_b200.storage._atomic import publish_text


def _circuit_hash(circuit_dict: dict) -> str:
    """Identity of a circuit: must equal the reference's hash (wal.py:17-22) so a work
    directory started by either implementation is recognised by the other."""
    canon = json.dumps(validate_circuit_dict(circuit_dict), sort_keys=True, default=str)
    return hashlib.sha256(canon.encode()).hexdigest()[:16]


class WAL:
    def __init__(self, path: str | Path, circuit_dict: dict | None = None):
        self.path = Path(path)
        self.path.parent.mkdir(parents=True, exist_ok=True)
        self._chash = _circuit_hash(circuit_dict) if circuit_dict else None
        if self.path.exists():
            self._data = json.loads(self.path.read_text())
            seen = self._data.get("circuit_hash")
            if self._chash and seen and seen != self._chash:
                raise ValueError("WAL circuit hash mismatch — different circuit? "
                                 f"WAL={seen} vs new={self._chash}")
        else:
            self._data = {"circuit_hash": self._chash or "", "committed_buf": "a", "done_steps": 0}
            self._flush()

    def _flush(self) -> None:
        publish_text(self.path, json.dumps(self._data, indent=2))

    @property
    def committed_buf(self) -> str:
        return self._data.get("committed_buf", "a")

    @property
    def done_steps(self) -> int:
        return self._data.get("done_steps", 0)

    def commit_step(self, step_idx: int, new_buf: str) -> None:
        self._data.update(committed_buf=new_buf, done_steps=step_idx + 1)
        self._flush()

    def close(self) -> None:  # API compatibility; nothing is held open
        pass
