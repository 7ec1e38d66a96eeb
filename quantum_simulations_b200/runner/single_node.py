"""Runner: circuit dict -> final state buffer, same surface as the reference's
``wenbo_engine.runner.single_node`` (run :78-138, collect_state :326-346).

What changes underneath: the reference streams every chunk FILE through
NumPy once per step and fsyncs it back; here the state lives in HBM for the whole run, a
"step" is a compiled sequence of fused on-chip passes (circuit/passes.py), and the
chunk/manifest/WAL machinery is kept for what it is still needed for — durable checkpoints
and the final hand-off (`run` still returns the path of a buffer directory that
`collect_state` reads).  Double-buffer semantics are unchanged: a checkpoint is written to
the buffer that is NOT the committed one, then manifest, then WAL commit, so a crash at any
point leaves the committed buffer intact and `run` resumes from ``wal.done_steps``.

kwargs keep the reference's names and meaning:
  chunk_size      amplitudes per checkpoint chunk file (must divide 2^n; clamped to 2^n)
  kernel          only "cuda" exists here ("scalar"/"batched" are the reference's CPU kernels)
  use_fusion      False: one compiled program per circuit level (the reference's one I/O pass
                  per level); True: levels are batched and compiled together (batch_levels)
  use_staging     on one GPU every qubit is local: atlas_stages(cd, k=n) is the batched circuit with the identity
                  mapping, which is written to qubit_mapping.json as the reference does (single_node.py:130-134);
                  sharded runs take their stages from atlas_stages in runner.multi_gpu.run(use_staging=True)
new kwargs: dtype ("complex128" default | "complex64"), device, checkpoint_every (steps
between durable checkpoints; 0 = only the final state), tile_bits / low_bits (pass compiler).
"""
from __future__ import annotations

import json
import math
import os
import shutil
from pathlib import Path

import numpy as np

from quantum_simulations_b200.circuit.fusion import batch_levels, _compile_ops
from quantum_simulations_b200.circuit.io import levelize, validate_circuit_dict
from quantum_simulations_b200.circuit.passes import PassCompiler, REG_BITS
from quantum_simulations_b200.storage.block_store import chunk_filename, read_chunk, write_chunk_atomic
from quantum_simulations_b200.storage.manifest import Manifest, read_manifest, write_manifest_atomic
from quantum_simulations_b200.wal.fencing import FencingLock
from quantum_simulations_b200.wal.wal import WAL


def _buf_dir(work: Path, buf: str) -> Path:
    return work / f"state_{buf}"


def _other(buf: str) -> str:
    return "a" if buf == "b" else "b"


def _crash_after() -> int | None:
    """Fault injection, same env var as the reference (single_node.py:61-63): die after this
    many chunk files of a checkpoint have been written."""
    val = os.environ.get("WE_CRASH_AFTER_CHUNK")
    return None if val is None else int(val)


def _crash_at_checkpoint() -> int:
    """Which checkpoint of the run WE_CRASH_AFTER_CHUNK applies to (extension: the reference always
    dies in its first step; resuming from a COMMITTED checkpoint needs a later one)."""
    return int(os.environ.get("WE_CRASH_AT_CHECKPOINT", "0"))


def _wipe_buf(buf_dir: Path) -> None:
    shutil.rmtree(buf_dir / "chunks", ignore_errors=True)
    (buf_dir / "manifest.json").unlink(missing_ok=True)


def build_steps(cd: dict, k: int, use_fusion: bool) -> list[dict]:
    """Step IR exactly as the reference builds it (single_node.py:108-121)."""
    levels = levelize(cd)
    if use_fusion:
        return batch_levels(levels, k)
    steps = []
    for lv in levels:
        if lv:
            loc, nonloc = _compile_ops(lv, k)
            steps.append({"local_ops": loc, "nonlocal_ops": nonloc})
    return steps


def execute_ops(state, ops, compiler: PassCompiler | None) -> None:
    """Apply a step's op list on the device: fused passes when the shard is large enough for
    the pass kernel, the per-gate kernels otherwise (both are CUDA; there is no CPU path)."""
    if not ops:
        return
    if compiler is None:
        for qs, U in ops:
            state.apply_op(qs, U)
    else:
        state.run_program(compiler.compile(ops))


def run(
    circuit_dict: dict,
    work_dir: str | Path,
    chunk_size: int = 1 << 20,
    kernel: str = "cuda",
    use_wal: bool = True,
    use_fencing: bool = False,
    use_fusion: bool = False,
    use_staging: bool = False,
    staging_method: str = "heuristic",
    dtype: str = "complex128",
    device: int = 0,
    checkpoint_every: int = 0,
    tile_bits: int | None = None,
    low_bits: int | None = None,
) -> Path:
    """Run the circuit on the GPU; returns the path of the final state buffer."""
    if kernel != "cuda":
        raise ValueError(f"kernel={kernel!r}: this engine only has kernel='cuda' "
                         "(the reference's 'scalar'/'batched' are CPU NumPy kernels)")
    if staging_method not in ("heuristic", "greedy", "ilp"):
        raise ValueError(f"unknown staging method: {staging_method!r}")
    cd = validate_circuit_dict(circuit_dict)
    n = cd["number_of_qubits"]
    N = 1 << n
    chunk_size = min(chunk_size, N)
    if N % chunk_size != 0:
        raise ValueError("2^n must be divisible by chunk_size")
    np_dtype = np.dtype(dtype)
    work = Path(work_dir)
    steps = build_steps(cd, n, use_fusion)          # one device: k = n, every op is local
    if use_staging:                                 # reference single_node.py:108-134 with k = n: nothing to stage
        from quantum_simulations_b200.circuit.staging import atlas_stages
        from quantum_simulations_b200.storage._atomic import publish_text
        steps, log_to_phys = atlas_stages(cd, n, method=staging_method)
        work.mkdir(parents=True, exist_ok=True)
        publish_text(work / "qubit_mapping.json", json.dumps(list(log_to_phys)))

    fence = FencingLock(work) if use_fencing else None
    if fence:
        fence.acquire()
    try:
        return _run_inner(cd, n, chunk_size, work, steps, use_wal, np_dtype, device,
                          checkpoint_every, tile_bits, low_bits)
    finally:
        if fence:
            fence.release()


def _run_inner(cd, n, chunk_size, work, steps, use_wal, np_dtype, device, checkpoint_every,
               tile_bits, low_bits) -> Path:
    from quantum_simulations_b200.kernel.cuda import DeviceState

    wal = WAL(work / "wal.json", circuit_dict=cd) if use_wal else None
    start = wal.done_steps if wal else 0
    current = wal.committed_buf if wal else "a"
    man = Manifest(n_qubits=n, chunk_size=chunk_size, n_chunks=(1 << n) // chunk_size,
                   dtype=np_dtype.name, chunks=[chunk_filename(i) for i in range((1 << n) // chunk_size)])
    compiler = PassCompiler(n, dtype=np_dtype.name, tile_bits=tile_bits, low_bits=low_bits) \
        if n >= REG_BITS else None

    with DeviceState(n, np_dtype, device) as st:
        if start > 0:                                   # resume from the committed checkpoint
            _load_checkpoint(st, _buf_dir(work, current), np_dtype)
        else:
            st.init_zero()
        last_ckpt = start
        n_ckpt = 0
        for idx in range(start, len(steps)):
            step = steps[idx]
            if step["nonlocal_ops"]:
                raise NotImplementedError("non-local gate on a single-device run")
            execute_ops(st, step["local_ops"], compiler)
            final = idx == len(steps) - 1
            if final or (checkpoint_every and (idx + 1 - last_ckpt) >= checkpoint_every):
                # durable checkpoint into the buffer that is NOT the committed one
                dst = _other(current)
                _write_checkpoint(st, _buf_dir(work, dst), man, np_dtype, inject=(n_ckpt == _crash_at_checkpoint()))
                n_ckpt += 1
                current, last_ckpt = dst, idx + 1
                if wal:
                    wal.commit_step(idx, current)
        if not (_buf_dir(work, current) / "manifest.json").exists():
            _write_checkpoint(st, _buf_dir(work, current), man, np_dtype)   # empty circuit
    if wal:
        wal.close()
    return _buf_dir(work, current)


def _write_checkpoint(st, dst_dir: Path, man: Manifest, np_dtype, inject: bool = True) -> None:
    """Shard -> chunk files through two pinned staging buffers: the D2H copy of chunk c+1 is
    in flight while chunk c is written and fsynced (reference _writer thread, pipeline.py:73-82)."""
    from quantum_simulations_b200.storage.pinned import PinnedBuffer

    _wipe_buf(dst_dir)
    crash_after = _crash_after() if inject else None
    cs = man.chunk_size
    nbytes = cs * np_dtype.itemsize
    bufs = [PinnedBuffer(nbytes), PinnedBuffer(nbytes)]
    try:
        st._ck(st.lib.qsv_download_async(st._h, bufs[0].ptr, 0, cs))
        for c in range(man.n_chunks):
            st.sync()
            if c + 1 < man.n_chunks:
                st._ck(st.lib.qsv_download_async(st._h, bufs[(c + 1) & 1].ptr, (c + 1) * cs, cs))
            write_chunk_atomic(dst_dir / "chunks" / man.chunks[c], bufs[c & 1].array(np_dtype, cs), np_dtype)
            if crash_after is not None and c + 1 >= crash_after:
                os._exit(1)
        st.sync()
    finally:
        for b in bufs:
            b.free()
    write_manifest_atomic(dst_dir, man)


def _load_checkpoint(st, src_dir: Path, np_dtype) -> None:
    m = read_manifest(src_dir)
    for c, name in enumerate(m.chunks):
        data = read_chunk(src_dir / "chunks" / name, np.dtype(m.dtype))
        st.upload(data.astype(np_dtype, copy=False), c * m.chunk_size)


def collect_state(buf_path: str | Path, apply_permutation: bool = False,
                  work_dir: str | Path | None = None) -> np.ndarray:
    """All chunks of a buffer as one complex128 array (reference single_node.py:326-346)."""
    m = read_manifest(buf_path)
    root = Path(buf_path) / "chunks"
    state = np.concatenate([read_chunk(root / c, np.dtype(m.dtype)) for c in m.chunks]).astype(np.complex128)
    if apply_permutation and work_dir is not None:
        mp = Path(work_dir) / "qubit_mapping.json"
        if mp.exists():
            from quantum_simulations_b200.circuit.staging import permute_state
            state = permute_state(state, json.loads(mp.read_text()))
    return state
