"""Crash recovery (reference wenbo_engine/wal/recovery.py:16-34): with double-buffered
checkpoints recovery is "run again" — the runner reads wal.json, reloads the committed
checkpoint into HBM, wipes the half-written other buffer and continues from ``done_steps``."""
from __future__ import annotations

from pathlib import Path


def recover(circuit_dict: dict, work_dir: str | Path, chunk_size: int = 1 << 20, **run_kw) -> Path | None:
    """Path of the final state buffer, or None if there is nothing to recover (no WAL file)."""
    if not (Path(work_dir) / "wal.json").exists():
        return None
    from quantum_simulations_b200.runner.single_node import run
    return run(circuit_dict, work_dir, chunk_size=chunk_size, use_wal=True, **run_kw)
