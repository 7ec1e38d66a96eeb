"""Race check of a run-time specialised pass kernel ON THE CPU (test infrastructure).

The generated CUDA source is built like tests/jit_host_run.py builds it (host prelude: one OS thread per CUDA thread,
mbarriers / named barriers as C++ atomics with acquire / release ordering, shared memory and the state as plain memory),
but as a stand-alone program with ``g++ -fsanitize=thread``.  ThreadSanitizer then reports every access to the ring
buffers or to the state that the kernel's own synchronisation (full / empty mbarriers, group barriers, __syncwarp) does
not order — a consumer reading a buffer before its fill arrived, a producer refilling a buffer a consumer still reads, a
round exchange that crosses warps behind a warp-level barrier, two threads storing to one address.  It is the host-model
counterpart of ``compute-sanitizer --tool racecheck``; it says nothing about the hardware memory model itself."""
from __future__ import annotations

import hashlib
import os
import struct
import subprocess
import tempfile
from pathlib import Path

import numpy as np

from tests.jit_host_run import CSRC, HOST_INC, kernel_coefs, kernel_source

_BUILD = Path(tempfile.gettempdir()) / "qsv_jit_tsan"


def tsan_available() -> bool:
    probe = _BUILD / "probe"
    _BUILD.mkdir(exist_ok=True)
    src = _BUILD / "probe.cpp"
    src.write_text("#include <thread>\nint main(){int x=0;std::thread t([&]{x=1;});t.join();return x-1;}\n")
    r = subprocess.run(["g++", "-fsanitize=thread", "-O1", "-pthread", "-o", str(probe), str(src)], capture_output=True, text=True)
    if r.returncode != 0:
        return False
    return subprocess.run([str(probe)], capture_output=True).returncode == 0


def race_check(step, n_local: int, dtype="complex128", grid: int = 1, tile_block: int = 0, tile_range=None,
               timeout: int = 600, mutate=None, stop_at_first: bool = False, sanitizer: str = "thread") -> tuple[int, str]:
    """(number of ThreadSanitizer reports, their text) for one launch of the pass's specialised kernel over a random
    state of 2^n_local amplitudes.  mutate(source) -> source: the positive controls of the test suite break the
    kernel's synchronisation on purpose and must be reported.
    sanitizer="address,undefined": the same program under AddressSanitizer + UBSan instead — the host-model counterpart of
    ``compute-sanitizer --tool memcheck``: the state is one heap block of exactly 2^n_local amplitudes and the ring is one
    heap block, so a global index outside the shard or a slot index outside the ring is reported."""
    import re
    src = kernel_source(step, dtype)
    if mutate is not None:
        src = mutate(src)
    src = "\n".join(ln for ln in src.splitlines() if "asm volatile" not in ln)
    call = "k_pass_jit((JV *)state, (const double2 *)tables, rank_bits, tile_begin, n_tiles, C, F)"
    full = src + f"\n#define JIT_HOST_CALL {call}\n#include \"jit_tsan_main.inc\"\n"
    key = hashlib.sha1((sanitizer + full + (HOST_INC / "jit_prelude.cuh").read_text() + (HOST_INC / "jit_tsan_main.inc").read_text()
                        + (HOST_INC / "jit_host_main.inc").read_text() + (CSRC / "pass_ops.cuh").read_text()).encode()).hexdigest()[:20]
    _BUILD.mkdir(exist_ok=True)
    exe = _BUILD / f"k_{key}"
    if not exe.exists():
        cpp = _BUILD / f"k_{key}.{os.getpid()}.cpp"
        tmp = _BUILD / f"k_{key}.{os.getpid()}"
        cpp.write_text(full)
        r = subprocess.run(["g++", f"-fsanitize={sanitizer}", "-fno-sanitize-recover=undefined", "-O1", "-g", "-std=c++17", "-pthread", "-Wno-unknown-pragmas", "-Wno-attributes",
                            f"-I{HOST_INC}", f"-I{CSRC}", "-o", str(tmp), str(cpp)], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ -fsanitize=thread failed on the generated kernel:\n" + r.stderr[-4000:])
        os.replace(tmp, exe)
        cpp.unlink()
    nt = int(re.search(r"__launch_bounds__\((\d+), 1\)", src).group(1))
    coefs = kernel_coefs(step, dtype).astype(np.float64)
    tables = np.ascontiguousarray(step.tables if step.tables is not None else np.zeros(0), dtype=np.complex128)
    n_tiles = (1 << n_local) >> 11
    tb, te = tile_range if tile_range is not None else (0, n_tiles)
    blob = _BUILD / f"in_{key}_{os.getpid()}.bin"
    blob.write_bytes(struct.pack("<8Q", n_local, grid, nt, tile_block, tb, te, len(coefs), len(tables)) + coefs.tobytes() + tables.tobytes())
    try:
        r = subprocess.run([str(exe), str(blob)], capture_output=True, text=True, timeout=timeout,
                           env=dict(os.environ, TSAN_OPTIONS=f"halt_on_error={int(stop_at_first)} exitcode=0 report_signal_unsafe=0",
                                    ASAN_OPTIONS="detect_leaks=0 exitcode=0"))
    finally:
        blob.unlink(missing_ok=True)
    reports = r.stderr.count("WARNING: ThreadSanitizer") + r.stderr.count("ERROR: AddressSanitizer") + r.stderr.count("runtime error:")
    if reports == 0 and "done rc=0" not in r.stdout:
        raise RuntimeError(f"race-check run failed (exit {r.returncode}):\n{r.stdout[-500:]}\n{r.stderr[-3000:]}")
    return reports, r.stderr[-6000:]
