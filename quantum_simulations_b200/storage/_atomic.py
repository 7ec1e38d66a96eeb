"""tmp + fsync + rename: the atomic-publish primitive the reference uses for chunks,
manifests, the WAL and the lock file (block_store.py:18-28, manifest.py:42-55,
wal.py:69-76, fencing.py:40-52)."""
from __future__ import annotations

import os
from pathlib import Path


def publish_bytes(path: str | Path, payload: bytes | memoryview) -> Path:
    dst = Path(path)
    dst.parent.mkdir(parents=True, exist_ok=True)
    tmp = dst.with_name(dst.name + ".tmp") if dst.suffix != ".bin" else dst.with_suffix(".tmp")
    fd = os.open(tmp, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
    try:
        view = memoryview(payload)
        done = 0
        while done < len(view):
            done += os.write(fd, view[done:done + (1 << 30)])
        os.fsync(fd)
    finally:
        os.close(fd)
    os.replace(tmp, dst)
    return dst


def publish_text(path: str | Path, text: str) -> Path:
    return publish_bytes(path, text.encode())
