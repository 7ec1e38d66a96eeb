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


def fuse_2q_blocks(ops: list[Op], only_diagonal: bool = True, tol: float = 0.0) -> list[Op]:
    """Merge runs of gates that stay inside one qubit PAIR into a single 4x4 unitary (the 2-qubit
    analogue of fuse_1q_ops, reference fusion.py:41-81).

    only_diagonal=True (default): a run is merged only if its product is exactly DIAGONAL.  That is
    the case front ends create when they decompose controlled phases into CNOT ladders
    (QASMBench: u1 / cx / u1 / cx / u1): the product of such a run has exact zeros off the diagonal
    (permutations and diagonals multiply without rounding), and the pass compiler then treats it as
    a phase gate instead of two controlled swaps.  Everything else is emitted unchanged, because a
    dense 4x4 block would cost the lifting form its structure (a general 1-qubit gate is 4.5 FP64 /
    amplitude, a dense 2-qubit block 16).  only_diagonal=False merges every run (dense blocks).
    tol > 0 also accepts runs whose product is diagonal up to `tol` (compiled ZZ interactions: three
    CNOTs and rotations by multiples of pi/2, exact only in exact arithmetic); the off-diagonal
    residue is dropped and the diagonal renormalised, an error of order tol per merged run."""
    out: list[Op] = []
    block_of: dict[int, int] = {}                 # qubit -> index into `blocks`
    blocks: list = []                             # [a, b, [(qs, U, U4)]] or None once emitted
    pending: dict[int, list] = {}                 # 1-qubit ops waiting for a pair run on their qubit
    I2 = np.eye(2, dtype=np.complex128)
    swap = np.eye(4)[[0, 2, 1, 3]]

    def diagonal_run(seq: list):
        """Longest prefix of `seq` (with a 2-qubit member) whose product is diagonal -> (length, diag) | None."""
        best, prod, has2 = None, np.eye(4, dtype=np.complex128), False
        for j, m in enumerate(seq):
            prod = m[2] @ prod
            has2 = has2 or len(m[0]) == 2
            if j > 0 and has2:
                off = prod - np.diag(np.diag(prod))
                if not np.any(off):
                    best = (j + 1, prod.copy())
                elif tol > 0 and np.abs(off).max() <= tol:
                    d = np.diag(prod)
                    best = (j + 1, np.diag(d / np.abs(d)))
        return best

    def emit_block(a: int, b: int, members: list, na: int = 0, nb: int = 0) -> None:
        if only_diagonal and (na or nb):
            # the block opens with 1-qubit gates that were waiting on its two qubits (members[:na] on a,
            # members[na:na+nb] on b; the two groups commute): a merged run may start anywhere in EACH
            # group, e.g. skip the Hadamard layer before a compiled ZZ block on one qubit only
            la, lb, rest = members[:na], members[na:na + nb], members[na + nb:]
            found = None
            for skip in range(na + nb + 1):                      # fewest skipped gates first
                for sa in range(max(0, skip - nb), min(na, skip) + 1):
                    sb = skip - sa
                    hit = diagonal_run(la[sa:] + lb[sb:] + rest)
                    if hit is not None and hit[0] > (na - sa) + (nb - sb):
                        found = (sa, sb, hit)
                        break
                if found:
                    break
            if found:
                sa, sb, (length, diag) = found
                out.extend((m[0], m[1]) for m in la[:sa] + lb[:sb])
                out.append(([a, b], diag))
                members = (la[sa:] + lb[sb:] + rest)[length:]
        if not only_diagonal:
            U = np.eye(4, dtype=np.complex128)
            for _, _, U4 in members:
                U = U4 @ U
            out.append(([a, b], U)) if any(len(q) == 2 for q, _, _ in members) else out.extend((q, u) for q, u, _ in members)
            return
        # 1-qubit gates after the block's last 2-qubit gate that do not end up in a merged run go back to
        # `pending`: they may LEAD the next pair run on their qubit (compiled ZZ / CPHASE blocks open with
        # 1-qubit gates, and a gate on a qubit of an open block joins that block first)
        last2 = max((k for k, m in enumerate(members) if len(m[0]) == 2), default=-1)
        i = 0
        while i < len(members):
            best, prod = None, np.eye(4, dtype=np.complex128)
            for j in range(i, len(members)):
                prod = members[j][2] @ prod
                if j > i and any(len(members[k][0]) == 2 for k in range(i, j + 1)):
                    off = prod - np.diag(np.diag(prod))
                    if not np.any(off):
                        best = (j, prod.copy())
                    elif tol > 0 and np.abs(off).max() <= tol:
                        d = np.diag(prod)
                        best = (j, np.diag(d / np.abs(d)))
            if best is None:
                if i > last2:
                    pending.setdefault(members[i][0][0], []).append((members[i][0], members[i][1]))
                else:
                    out.append((members[i][0], members[i][1]))
                i += 1
            else:
                out.append(([a, b], best[1]))
                i = best[0] + 1

    def close(q: int) -> None:
        i = block_of.pop(q, None)
        if i is None or blocks[i] is None:
            return
        a, b, members, na, nb = blocks[i]
        blocks[i] = None
        block_of.pop(a, None)
        block_of.pop(b, None)
        emit_block(a, b, members, na, nb)

    for qs, U in ops:
        U = np.asarray(U, dtype=np.complex128)
        if len(qs) == 1:
            q = qs[0]
            i = block_of.get(q)
            if i is not None and blocks[i] is not None:
                a, b, members = blocks[i][:3]
                members.append((list(qs), U, np.kron(U, I2) if q == a else np.kron(I2, U)))
            else:
                pending.setdefault(q, []).append((list(qs), U))     # may lead a pair run that starts later
            continue
        a, b = qs
        ia, ib = block_of.get(a), block_of.get(b)
        if ia is not None and ia == ib and blocks[ia] is not None:
            x, y, members = blocks[ia][:3]
            members.append((list(qs), U, U if (x, y) == (a, b) else swap @ U @ swap))
            continue
        close(a)
        close(b)
        lead_a = [(q_, u_, np.kron(u_, I2)) for q_, u_ in pending.pop(a, [])]
        lead_b = [(q_, u_, np.kron(I2, u_)) for q_, u_ in pending.pop(b, [])]
        blocks.append([a, b, lead_a + lead_b + [(list(qs), U, U)], len(lead_a), len(lead_b)])
        block_of[a] = block_of[b] = len(blocks) - 1
    for blk in blocks:
        if blk is not None:
            emit_block(*blk)
    for q_ops in pending.values():
        out.extend(q_ops)
    return out
