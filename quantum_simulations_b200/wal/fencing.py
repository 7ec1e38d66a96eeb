"""run.lock fencing (reference wal/fencing.py:16-80): one live writer per work dir."""
from __future__ import annotations

import json
import os
import platform
import time
from pathlib import Path

from quantum_simulations_b200.storage._atomic import publish_text

_STALE_AFTER_S = 86400  # foreign-host locks older than a day are treated as dead


class FencingLock:
    def __init__(self, work_dir: str | Path):
        self.lock_path = Path(work_dir) / "run.lock"
        self.lock_path.parent.mkdir(parents=True, exist_ok=True)

    def _holder(self) -> dict | None:
        try:
            return json.loads(self.lock_path.read_text())
        except (OSError, json.JSONDecodeError):
            return None

    @staticmethod
    def _is_alive(holder: dict) -> bool:
        if holder.get("host") != platform.node():
            return time.time() - holder.get("ts", 0) < _STALE_AFTER_S
        try:
            os.kill(holder["pid"], 0)
        except (OSError, ProcessLookupError):
            return False
        return True

    def acquire(self, force: bool = False) -> None:
        if not force and self.lock_path.exists():
            holder = self._holder()
            if holder and self._is_alive(holder):
                raise RuntimeError(
                    f"Work directory locked by PID {holder['pid']} on {holder['host']} "
                    f"since {holder.get('time', '?')}. If stale, delete {self.lock_path} "
                    "or use force=True.")
        publish_text(self.lock_path, json.dumps({
            "pid": os.getpid(), "host": platform.node(),
            "time": time.strftime("%Y-%m-%d %H:%M:%S"), "ts": time.time()}))

    def release(self) -> None:
        self.lock_path.unlink(missing_ok=True)

    def __enter__(self):
        self.acquire()
        return self

    def __exit__(self, *exc):
        self.release()
