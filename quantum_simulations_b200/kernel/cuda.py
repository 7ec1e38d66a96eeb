"""kernel="cuda": the GPU operator face of the runner.

Two faces, both over the C ABI in include/qsv.h (ctypes, no torch):

* ``apply_1q(chunk, qubit, U)`` / ``apply_2q(chunk, qa, qb, U)`` — drop-in for the pair of
  callables the reference runner threads down as ``a1, a2``
  (wenbo_engine/runner/single_node.py:103-106; kernel/cpu_scalar.py:21-47): in place on a
  host ndarray, ``NotImplementedError("... non-local ...")`` for a qubit >= log2(len(chunk)).
  They round-trip the chunk over PCIe, so they are for parity tests and small chunks.
* ``DeviceState`` — the state (or one shard of it) resident in HBM; this is what
  ``runner.single_node.run(kernel="cuda")`` drives.

There is no CPU fallback: without libqsv.so or without a CUDA device every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from quantum_simulations_b200 import _lib as L
from quantum_simulations_b200.circuit.passes import PassStep, Dense1QStep, Dense2QStep, Program, lower_op, Dense2Q

_DTYPES = {"complex64": L.QSV_C64, "complex128": L.QSV_C128}


def _mat(U, dim: int):
    u = np.ascontiguousarray(U, dtype=np.complex128)
    if u.shape != (dim, dim):
        raise ValueError(f"expected a {dim}x{dim} matrix, got {u.shape}")
    return u, u.ctypes.data_as(C.POINTER(C.c_double))


_MARGINAL_BITS = 10          # bins of a marginal live in shared memory up to 2^10 (csrc/sample.cuh k_marginal)


def pad_marginal_qubits(qubits: list, n_local: int, bits: int = _MARGINAL_BITS) -> list:
    """`qubits` followed by the lowest local qubits not among them, up to `bits` in total (never fewer than asked)."""
    have = set(qubits)
    extra = [q for q in range(n_local) if q not in have][: max(0, bits - len(qubits))]
    return list(qubits) + extra


def fold_marginal(wide: np.ndarray, asked: int) -> np.ndarray:
    """sum a marginal over the padding bits: outcome bit k of `wide` is qubit k of the widened list, the first
    `asked` of which are the caller's"""
    return np.ascontiguousarray(wide.reshape(-1, 1 << asked).sum(axis=0))


class DeviceState:
    """2^n_local amplitudes of an n-qubit state in HBM (shard `rank` of `world`)."""

    def __init__(self, n_qubits: int, dtype="complex128", device: int = 0, rank: int = 0, world: int = 1):
        self.lib = L.load()
        self.dtype = np.dtype(dtype)
        if self.dtype.name not in _DTYPES:
            raise ValueError(f"unsupported dtype {dtype}")
        self.n_qubits, self.rank, self.world, self.device = n_qubits, rank, world, device
        self.n_local = n_qubits - int(math.log2(world))
        self.n_amps = 1 << self.n_local
        self._h = C.c_void_p()
        rc = self.lib.qsv_create(C.byref(self._h), n_qubits, _DTYPES[self.dtype.name], device, rank, world)
        if rc:
            msg = self.lib.qsv_last_error(None)
            raise L.QsvError(rc, msg.decode() if msg else "?")
        self._programs: list = []

    # ---- lifecycle ----
    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            for p in self._programs:
                self.lib.qsv_program_destroy(self._h, p)
            self._programs.clear()
            self.lib.qsv_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int) -> None:
        if rc:
            msg = self.lib.qsv_last_error(self._h)
            text = msg.decode(errors="replace") if msg else "?"
            if rc == L.QSV_ENONLOCAL:
                raise NotImplementedError(text)
            raise L.QsvError(rc, text)

    def sync(self) -> None:
        self._ck(self.lib.qsv_sync(self._h))

    # ---- state I/O ----
    def init_zero(self) -> None:
        self._ck(self.lib.qsv_init_zero(self._h))

    def init_basis(self, index: int) -> None:
        self._ck(self.lib.qsv_init_basis(self._h, index))

    def upload(self, host: np.ndarray, offset: int = 0) -> None:
        a = np.ascontiguousarray(host, dtype=self.dtype)
        self._ck(self.lib.qsv_upload(self._h, a.ctypes.data, offset, a.size))

    def download(self, out: np.ndarray | None = None, offset: int = 0, count: int | None = None) -> np.ndarray:
        count = self.n_amps - offset if count is None else count
        if out is None:
            out = np.empty(count, dtype=self.dtype)
        if out.dtype != self.dtype or not out.flags.c_contiguous or out.size < count:
            raise ValueError("download buffer must be a contiguous array of the state's dtype")
        self._ck(self.lib.qsv_download(self._h, out.ctypes.data, offset, count))
        return out

    def stream_to_host(self, sink, chunk_amps: int = 1 << 24, offset: int = 0, count: int | None = None) -> int:
        """Hand the shard to `sink(view, first_amp)` piece by piece through TWO pinned staging buffers (the copy
        of piece c+1 runs while the sink works on piece c) — the way a state that does not fit in host memory
        leaves the device (checkpoint writers, reductions).  Views are only valid inside the call.  Returns
        the bytes copied."""
        from quantum_simulations_b200.storage.pinned import PinnedBuffer
        count = self.n_amps - offset if count is None else count
        chunk_amps = max(1, min(chunk_amps, count))
        bufs = [PinnedBuffer(chunk_amps * self.dtype.itemsize) for _ in range(2)]
        try:
            pieces = [(o, min(chunk_amps, offset + count - o)) for o in range(offset, offset + count, chunk_amps)]
            if pieces:
                self._ck(self.lib.qsv_download_async(self._h, bufs[0].ptr, pieces[0][0], pieces[0][1]))
            for c, (o, m) in enumerate(pieces):
                self.sync()
                if c + 1 < len(pieces):
                    self._ck(self.lib.qsv_download_async(self._h, bufs[(c + 1) & 1].ptr, pieces[c + 1][0], pieces[c + 1][1]))
                sink(bufs[c & 1].array(self.dtype, m), o)
            self.sync()
        finally:
            for b in bufs:
                b.free()
        return count * self.dtype.itemsize

    def device_ptr(self) -> tuple[int, int, int]:
        p, n, s = C.c_void_p(), C.c_size_t(), C.c_void_p()
        self._ck(self.lib.qsv_device_ptr(self._h, C.byref(p), C.byref(n), C.byref(s)))
        return p.value, n.value, s.value or 0

    # ---- per-gate operators ----
    def apply_1q(self, q: int, U) -> None:
        _, p = _mat(U, 2)
        self._ck(self.lib.qsv_apply_1q(self._h, q, p))

    def apply_2q(self, qa: int, qb: int, U) -> None:
        _, p = _mat(U, 4)
        self._ck(self.lib.qsv_apply_2q(self._h, qa, qb, p))

    def apply_ctrl_1q(self, ctrl: int, tgt: int, U) -> None:
        _, p = _mat(U, 2)
        self._ck(self.lib.qsv_apply_ctrl_1q(self._h, ctrl, tgt, p))

    def apply_diag(self, qubits, phases) -> None:
        qs = (C.c_int * len(qubits))(*qubits)
        ph = np.ascontiguousarray(phases, dtype=np.complex128)
        if ph.size != 1 << len(qubits):
            raise ValueError("need 2^nq phases")
        self._ck(self.lib.qsv_apply_diag(self._h, len(qubits), qs, ph.ctypes.data_as(C.POINTER(C.c_double))))

    def apply_kq(self, qubits, U) -> None:
        k = len(qubits)
        qs = (C.c_int * k)(*qubits)
        _, p = _mat(U, 1 << k)
        self._ck(self.lib.qsv_apply_kq(self._h, k, qs, p))

    def apply_op(self, qubits, U) -> None:
        """One step-IR op through the per-gate kernels, picking the specialised kernel from
        the matrix structure (diagonal / controlled / dense)."""
        qubits = list(qubits)
        if len(qubits) == 1:
            u = np.asarray(U)
            if u[0, 1] == 0 and u[1, 0] == 0:
                return self.apply_diag(qubits, np.diag(u))
            return self.apply_1q(qubits[0], U)
        u = np.asarray(U, dtype=np.complex128)
        if not np.any(u - np.diag(np.diag(u))):
            return self.apply_diag(qubits, np.diag(u))
        low = lower_op(qubits, u)
        if len(low) == 1 and not isinstance(low[0], (tuple, Dense2Q)) and low[0].target is not None \
                and len(low[0].ctrls) == 1:
            (c,), t = low[0].ctrls, low[0].target
            sub = u[2:, 2:] if c == qubits[0] else u[np.ix_([1, 3], [1, 3])]
            return self.apply_ctrl_1q(c, t, sub)
        return self.apply_2q(qubits[0], qubits[1], u)

    # ---- fused passes ----
    def apply_pass(self, step: PassStep) -> None:
        tab = None if step.tables is None else step.tables.ctypes.data_as(C.POINTER(C.c_double))
        self._ck(self.lib.qsv_apply_pass(self._h, C.byref(step.desc), step.ops, tab))

    def set_option(self, option: int, value: int) -> None:
        self._ck(self.lib.qsv_set_option(self._h, option, value))

    def run_program(self, prog: Program, jit: bool | None = None) -> None:
        """Execute a compiled program (passes + stand-alone dense 2q kernels) in order.

        jit=None/True: maximal runs of passes go through qsv_program_create, which specialises
        them at run time (csrc/jit.cuh); jit=False: every pass is interpreted (one-shot
        qsv_apply_pass)."""
        if prog.n_local != self.n_local or prog.dtype != self.dtype.name:
            raise ValueError("program was compiled for a different shard shape / dtype")
        if jit is None or jit:
            run: list = []
            for step in list(prog.steps) + [None]:
                if isinstance(step, PassStep):
                    run.append(step)
                    continue
                if run:
                    h = self.upload_steps(run)
                    self.replay(h)
                    self.release_program(h)
                    run = []
                if isinstance(step, Dense2QStep):
                    self.apply_2q(step.qa_pos, step.qb_pos, step.U)
                elif isinstance(step, Dense1QStep):
                    self.apply_1q(step.q_pos, step.U)
                elif step is not None:
                    raise TypeError(type(step))
            return
        for step in prog.steps:
            if isinstance(step, PassStep):
                self.apply_pass(step)
            elif isinstance(step, Dense2QStep):
                self.apply_2q(step.qa_pos, step.qb_pos, step.U)
            elif isinstance(step, Dense1QStep):
                self.apply_1q(step.q_pos, step.U)
            else:
                raise TypeError(type(step))

    def upload_program(self, prog: Program):
        """Device-resident copy of an all-pass program for replay (qsv_program_*)."""
        steps = prog.steps
        if not all(isinstance(s, PassStep) for s in steps):
            raise ValueError("only all-pass programs can be uploaded")
        return self.upload_steps(steps)

    def release_program(self, handle) -> None:
        self._ck(self.lib.qsv_program_destroy(self._h, handle))
        self._programs = [p for p in self._programs if p.value != handle.value]

    def upload_steps(self, steps):
        n = len(steps)
        passes = (L.QsvPass * max(n, 1))(*[s.desc for s in steps])
        total = sum(s.n_micro_ops for s in steps)
        ops = (L.QsvOp * max(total, 1))()
        k = 0
        for s in steps:
            for j in range(s.n_micro_ops):
                ops[k] = s.ops[j]
                k += 1
        tabs = [s.tables for s in steps if s.tables is not None]
        tab = np.ascontiguousarray(np.concatenate(tabs)) if tabs else None
        tab_p = None if tab is None else tab.ctypes.data_as(C.POINTER(C.c_double))
        out = C.c_void_p()
        self._ck(self.lib.qsv_program_create(self._h, passes, n, ops, tab_p, C.byref(out)))
        self._programs.append(out)
        return out

    def replay(self, handle) -> None:
        self._ck(self.lib.qsv_program_run(self._h, handle))

    # ---- reductions / timing ----
    def norm2(self) -> float:
        out = C.c_double()
        self._ck(self.lib.qsv_norm2(self._h, C.byref(out)))
        return out.value

    def probabilities(self, qubits) -> np.ndarray:
        """Marginal distribution of this shard over `qubits` (bit k of the outcome = qubits[k]).

        A marginal over FEW qubits is a histogram with few bins, and every thread of the kernel adds to one of them:
        2 bins cost 43-87 ms at 30 qubits where 1024 bins cost 5.7 ms (profiles/r02/kernel_table_n30_c128.txt).  So
        the request is widened with the lowest local qubits that are not in it (neighbouring lanes then hit
        different bins) up to 10 qubits = 1024 bins in shared memory, and the extra bits are summed away here."""
        qubits = list(qubits)
        asked = len(qubits)
        wide = pad_marginal_qubits(qubits, self.n_local)
        qs = (C.c_int * max(len(wide), 1))(*wide)
        out = np.zeros(1 << len(wide), dtype=np.float64)
        self._ck(self.lib.qsv_probabilities(self._h, len(wide), qs, out.ctypes.data_as(C.POINTER(C.c_double))))
        return fold_marginal(out, asked)

    def expect_z(self, qubits) -> float:
        """<Z_q1 Z_q2 ...> contribution of this shard."""
        mask = 0
        for q in qubits:
            mask |= 1 << q
        out = C.c_double()
        self._ck(self.lib.qsv_expect_z(self._h, mask, C.byref(out)))
        return out.value

    def project(self, qubit: int, outcome: int, renormalise: bool = True) -> float:
        """Partial measurement with a known outcome (HiSVSIM ``StateVector::project``,
        hisvsim_repo/state_vector.hpp:829-922): keep the amplitudes whose bit `qubit` equals `outcome`,
        zero the others and (optionally) renormalise.  Returns the probability of the outcome.
        One branch-free diagonal sweep (qsv_apply_diag with the table (s, 0) or (0, s))."""
        if outcome not in (0, 1):
            raise ValueError("outcome must be 0 or 1")
        p = float(self.probabilities([qubit])[outcome])
        if self.world > 1:
            raise NotImplementedError("project on a sharded state: sum the probabilities over the shards first "
                                      "(ShardedSimulator.project)")
        if p <= 0.0:
            raise ValueError(f"outcome {outcome} of qubit {qubit} has probability 0")
        s = 1.0 / np.sqrt(p) if renormalise else 1.0
        self.apply_diag([qubit], [s, 0.0] if outcome == 0 else [0.0, s])
        return p

    def sample(self, seed: int, shots: int) -> np.ndarray:
        """Measurement samples (basis-state indices, ascending), bit-exact with
        oracle/ref_dense.py::sample_indices: u = sort(default_rng(seed).random(shots))."""
        u = np.sort(np.random.default_rng(seed).random(shots))
        out = np.empty(shots, dtype=np.uint64)
        self._ck(self.lib.qsv_sample(self._h, seed, shots, u.ctypes.data_as(C.POINTER(C.c_double)),
                                     out.ctypes.data_as(C.POINTER(C.c_uint64))))
        return out

    def timing(self, on: bool) -> None:
        self._ck(self.lib.qsv_timing_enable(self._h, int(on)))

    def timer_start(self) -> None:
        self._ck(self.lib.qsv_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        self._ck(self.lib.qsv_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def take_timings(self, cap: int = 65536) -> list[tuple[float, int, int]]:
        buf = (L.QsvTiming * cap)()
        n = C.c_int()
        self._ck(self.lib.qsv_get_timings(self._h, buf, cap, C.byref(n)))
        return [(buf[i].ms, buf[i].kind, buf[i].pass_index) for i in range(min(n.value, cap))]


# ------------------------------------------------------------ a1 / a2 drop-in callables
def check_local(qubit: int, chunk_len: int) -> None:
    k = int(math.log2(chunk_len))
    if qubit >= k:
        raise NotImplementedError(
            f"qubit {qubit} >= log2(chunk_size)={k}: non-local gate requires layout/collect step")


def _on_device(chunk: np.ndarray, fn) -> None:
    if chunk.dtype not in (np.complex64, np.complex128) or chunk.ndim != 1:
        raise ValueError("chunk must be a 1-D complex64/complex128 array")
    n = int(math.log2(len(chunk)))
    if 1 << n != len(chunk):
        raise ValueError("chunk length must be a power of two")
    with DeviceState(n, chunk.dtype) as st:
        st.upload(chunk)
        fn(st)
        if chunk.flags.c_contiguous:
            st.download(chunk)
        else:
            chunk[:] = st.download()


def apply_1q(chunk: np.ndarray, qubit: int, U: np.ndarray) -> None:
    check_local(qubit, len(chunk))
    _on_device(chunk, lambda st: st.apply_1q(qubit, U))


def apply_2q(chunk: np.ndarray, qa: int, qb: int, U: np.ndarray) -> None:
    check_local(qa, len(chunk))
    check_local(qb, len(chunk))
    _on_device(chunk, lambda st: st.apply_2q(qa, qb, U))
