"""Level batching and 1-qubit gate fusion on the step IR.

Host-side mirror of wenbo_engine/circuit/fusion.py:22-165.  The step IR is the one
the reference runner consumes (fusion.py:91-96):

    {"local_ops": [(phys_qubits, U), ...], "nonlocal_ops": [(phys_qubits, U), ...],
     "level_indices": [...]}

where an op is *local* when all its qubits are < k = log2(chunk_size).  On the GPU a
"chunk" is one device shard, so k = n - log2(n_devices); on one B200 every op is local
and the whole circuit collapses to a single step, which the pass compiler
(`quantum_simulations_b200.circuit.passes`) then cuts into on-chip tile passes.
"""
from __future__ import annotations

import numpy as np

from quantum_simulations_b200.kernel import gates as gmod

Op = tuple  # (list[int] qubits, np.ndarray U)


def _compile_ops(level_gates: list[dict], k: int) -> tuple[list[Op], list[Op]]:
    """Gate dicts of one level -> (local ops, non-local ops) (reference fusion.py:22-36)."""
    buckets: tuple[list[Op], list[Op]] = ([], [])
    for g in level_gates:
        op = (g["qubits"], gmod.gate_matrix(g["gate"], g["params"]))
        buckets[0 if max(g["qubits"]) < k else 1].append(op)
    return buckets


def fuse_1q_ops(ops: list[Op]) -> list[Op]:
    """Pre-multiply runs of 1-qubit gates that hit the same qubit (reference :41-81).

    A run on qubit q is closed by any 2-qubit gate touching q (the accumulated matrix
    is emitted just before that gate).  Runs still open at the end are emitted in
    ascending qubit order.  Composition is ``new @ old``.
    """
    if not ops:
        return ops
    open_runs: dict[int, np.ndarray] = {}
    out: list[Op] = []
    for qubits, U in ops:
        if len(qubits) == 1:
            q = qubits[0]
            open_runs[q] = U @ open_runs[q] if q in open_runs else U.copy()
            continue
        for q in qubits:
            if q in open_runs:
                out.append(([q], open_runs.pop(q)))
        out.append((qubits, U))
    out.extend(([q], open_runs[q]) for q in sorted(open_runs))
    return out


def batch_levels(levels: list[list[dict]], k: int) -> list[dict]:
    """Merge consecutive all-local levels into one step (reference fusion.py:86-142).

    A level that contains any non-local op is emitted alone (its local ops unfused,
    exactly like the reference) and closes the batch before it.
    """
    steps: list[dict] = []
    acc_ops: list[Op] = []
    acc_idx: list[int] = []

    def close_batch() -> None:
        if acc_ops:
            steps.append({"local_ops": fuse_1q_ops(list(acc_ops)), "nonlocal_ops": [],
                          "level_indices": list(acc_idx)})
            acc_ops.clear()
            acc_idx.clear()

    for idx, gates in enumerate(levels):
        if not gates:
            continue
        loc, nonloc = _compile_ops(gates, k)
        if nonloc:
            close_batch()
            steps.append({"local_ops": loc, "nonlocal_ops": nonloc, "level_indices": [idx]})
        else:
            acc_ops.extend(loc)
            acc_idx.append(idx)
    close_batch()
    return steps


def fusion_stats(levels: list[list[dict]], k: int) -> dict:
    """What batching buys, for benches (reference fusion.py:145-165)."""
    steps = batch_levels(levels, k)
    n_levels = sum(1 for lv in levels if lv)
    n_steps = len(steps)
    saved = (1 - n_steps / max(n_levels, 1)) * 100
    return {
        "original_levels": n_levels,
        "fused_passes": n_steps,
        "local_only_passes": sum(1 for s in steps if not s["nonlocal_ops"]),
        "io_reduction": f"{n_levels}→{n_steps} ({saved:.0f}% fewer)",
        "ops_before": sum(len(lv) for lv in levels),
        "ops_after": sum(len(s["local_ops"]) + len(s["nonlocal_ops"]) for s in steps),
    }
