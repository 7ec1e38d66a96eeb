"""Run a run-time specialised pass kernel ON THE CPU (test infrastructure).

The CUDA source qsvjit::generate() emits is compiled with g++ against tests/jit_host/jit_prelude.cuh
(a host stand-in for the device prelude: one OS thread per CUDA thread, cp.async = 16-byte copy,
mbarriers = arrival counters, named barriers = real barriers; the gate bodies are the real
csrc/pass_ops.cuh) and called through ctypes.  This checks what only the generator decides — slot
codes, round exchanges, fold tables, op order, coefficient indices, load / store / scatter addressing,
the zero-input form — without a GPU.  It does not replace the GPU parity tests (no PTX, no hardware
memory model); it makes generator changes testable in `-m "not gpu"`."""
from __future__ import annotations

import ctypes as C
import hashlib
import subprocess
import tempfile
from pathlib import Path

import numpy as np

from quantum_simulations_b200 import _lib as L

ROOT = Path(__file__).resolve().parent.parent
HOST_INC = ROOT / "tests" / "jit_host"
CSRC = ROOT / "quantum_simulations_b200" / "csrc"
_BUILD = Path(tempfile.gettempdir()) / "qsv_jit_host"
_cache: dict = {}


def _dt(dtype) -> int:
    return L.QSV_C128 if np.dtype(dtype) == np.complex128 else L.QSV_C64


def kernel_source(step, dtype="complex128", scatter_bits=None) -> str:
    buf, need = C.create_string_buffer(1 << 20), C.c_size_t()
    lib = L.load()
    if scatter_bits is None:
        rc = lib.qsv_jit_source(C.byref(step.desc), step.ops, _dt(dtype), buf, len(buf), C.byref(need))
    else:
        lb = (C.c_int * len(scatter_bits))(*scatter_bits)
        rc = lib.qsv_jit_source_scatter(C.byref(step.desc), step.ops, _dt(dtype), len(scatter_bits), lb, buf, len(buf), C.byref(need))
    if rc != 0 or need.value >= len(buf):
        raise RuntimeError(f"pass is not eligible for specialisation (rc={rc})")
    return buf.value.decode()


def kernel_coefs(step, dtype="complex128") -> np.ndarray:
    need = C.c_size_t()
    out = (C.c_double * 512)()
    rc = L.load().qsv_jit_coefs(C.byref(step.desc), step.ops, _dt(dtype), out, 512, C.byref(need))
    assert rc == 0 and need.value <= 512
    return np.array(out[: max(need.value, 1)], dtype=np.float64)


def host_kernel(step, dtype="complex128", scatter_bits=None):
    """g++-compiled host build of the pass's specialised kernel -> (ctypes function, n_threads)."""
    src = kernel_source(step, dtype, scatter_bits)
    src = "\n".join(ln for ln in src.splitlines() if "asm volatile" not in ln)        # setmaxnreg hints
    call = "k_pass_jit((JV *)state, (const double2 *)tables, rank_bits, tile_begin, n_tiles, C, F" + \
           (", D)" if scatter_bits is not None else ")")
    full = src + f"\n#define JIT_HOST_CALL {call}\n#include \"jit_host_main.inc\"\n"
    key = hashlib.sha1((full + (HOST_INC / "jit_prelude.cuh").read_text() + (CSRC / "pass_ops.cuh").read_text()).encode()).hexdigest()[:20]
    if key not in _cache:
        _BUILD.mkdir(exist_ok=True)
        cpp, so = _BUILD / f"k_{key}.cpp", _BUILD / f"k_{key}.so"
        if not so.exists():
            import os
            cpp = _BUILD / f"k_{key}.{os.getpid()}.cpp"             # several test processes may build the same kernel
            tmp = _BUILD / f"k_{key}.{os.getpid()}.so"
            cpp.write_text(full)
            cmd = ["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-pthread", "-Wno-unknown-pragmas", "-Wno-attributes",
                   f"-I{HOST_INC}", f"-I{CSRC}", "-o", str(tmp), str(cpp)]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("g++ failed on the generated kernel:\n" + r.stderr[-4000:])
            os.replace(tmp, so)
            cpp.unlink()
        lib = C.CDLL(str(so))
        fn = lib.jit_host_run
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_ulonglong, C.c_uint, C.c_uint, C.c_void_p, C.POINTER(C.c_void_p),
                       C.c_ulonglong, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint), C.c_ulonglong, C.c_uint]
        import re
        nt = int(re.search(r"__launch_bounds__\((\d+), 1\)", src).group(1))
        _cache[key] = (fn, nt)
    return _cache[key]


def run_pass_on_host(step, shard: np.ndarray, n_local: int, rank: int = 0, grid: int = 2, tile_range=None,
                     scatter=None, fix=None, tile_block: int = 0) -> None:
    """Apply the pass to `shard` (in place) with the host build of its specialised kernel.
    scatter = (local_bits, targets, keep): the scatter variant; targets[x] = array that receives the
    amplitudes whose swapped local bits equal x (the shard itself is only read)."""
    dtype = shard.dtype
    fn, nt = host_kernel(step, dtype, None if scatter is None else list(scatter[0]))
    coefs = kernel_coefs(step, dtype).astype(np.float64 if dtype == np.complex128 else np.float32)
    tables = step.tables if step.tables is not None else np.zeros(1, dtype=np.complex128)
    tables = np.ascontiguousarray(tables, dtype=np.complex128)
    n_tiles = len(shard) >> 11
    tb, te = tile_range if tile_range is not None else (0, n_tiles)
    dst, keep = None, 0
    if scatter is not None:
        _, targets, keep = scatter
        dst = (C.c_void_p * 8)(*([t.ctypes.data for t in targets] + [None] * (8 - len(targets))))
    fpos, fval = fix if fix is not None else ((), 0)             # fix = (positions in tile-index space, value there)
    rc = fn(shard.ctypes.data, tables.ctypes.data, rank << n_local, tb, te, coefs.ctypes.data, dst, keep, grid, nt,
            len(fpos), (C.c_uint * 4)(*(list(fpos) + [0] * (4 - len(fpos)))), fval, tile_block)
    assert rc == 0
