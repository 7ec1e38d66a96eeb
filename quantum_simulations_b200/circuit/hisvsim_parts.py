"""HiSVSIM part files (v3_hisvsim_spark/hisvsim_repo/QASMBench/cluster/*/*_part_{smart,dfs,nat}):
the acyclic partition of a circuit's gate DAG into "parts" that HiSVSIM executes one after the
other, each on the qubits it touches (hisvsim_repo/svsim-mpi.hpp:123-173).

File format (one node per line):  ``<index> <name>_<node id> <part>``
    index >= 1   the index-th gate statement of the .qasm file, e.g. ``3 cx_4 2``
    index == 0   entry node of a qubit (``0 qr5 1``)            - ignored here
    ``..._exit_..``  exit node of a qubit                        - ignored here

`reorder_by_parts` turns such a partition into the gate order this engine runs: parts in a
topological order of the part DAG, gates inside a part in circuit order.  The reordering never
moves a gate across another gate on a shared qubit, so the state is unchanged; the pass compiler /
stage planner then sees the locality HiSVSIM's partitioner found (all gates of a part touch
at most `max(len(qubit set))` qubits)."""
from __future__ import annotations

import re

_NODE = re.compile(r"^\s*(\d+)\s+([A-Za-z][A-Za-z0-9]*)_(\d+)\s+(\d+)\s*$")


def read_part_file(text: str) -> list[int]:
    """Part id of gate statement 0, 1, 2, ... (file indices are 1-based)."""
    parts: dict[int, int] = {}
    for line in text.splitlines():
        if "_exit_" in line:
            continue
        m = _NODE.match(line)
        if not m or int(m.group(1)) == 0:
            continue
        idx = int(m.group(1)) - 1
        if idx in parts:
            raise ValueError(f"part file names gate {idx + 1} twice")
        parts[idx] = int(m.group(4))
    if sorted(parts) != list(range(len(parts))):
        raise ValueError("part file does not number the gates 1..G")
    return [parts[i] for i in range(len(parts))]


def reorder_by_parts(ops: list, part_of_op: list[int]) -> tuple[list, list[set[int]], list[int]]:
    """ops = [(qubits, U)], part_of_op[i] = part of ops[i]  ->  (reordered ops, qubit set of every
    part in execution order, part ids in execution order).  Raises ValueError when the partition
    is cyclic (two parts that each have to run before the other)."""
    if len(ops) != len(part_of_op):
        raise ValueError(f"{len(ops)} ops but {len(part_of_op)} part labels")
    ids = sorted(set(part_of_op))
    succ: dict[int, set[int]] = {p: set() for p in ids}
    indeg = {p: 0 for p in ids}
    last_part: dict[int, int] = {}
    for (qs, _), p in zip(ops, part_of_op):
        for q in qs:
            a = last_part.get(q)
            if a is not None and a != p and p not in succ[a]:
                succ[a].add(p)
                indeg[p] += 1
            last_part[q] = p
    ready = sorted(p for p in ids if indeg[p] == 0)
    order: list[int] = []
    while ready:
        p = ready.pop(0)
        order.append(p)
        for s_ in sorted(succ[p]):
            indeg[s_] -= 1
            if indeg[s_] == 0:
                ready.append(s_)
                ready.sort()
    if len(order) != len(ids):
        raise ValueError("the partition is cyclic: no order of the parts respects the gate dependencies")
    rank = {p: i for i, p in enumerate(order)}
    idx = sorted(range(len(ops)), key=lambda i: (rank[part_of_op[i]], i))
    qsets = [set() for _ in order]
    for i in idx:
        qsets[rank[part_of_op[i]]].update(ops[i][0])
    return [ops[i] for i in idx], qsets, order


def qasm_with_parts(qasm_text: str, part_text: str):
    """(n_qubits, reordered ops, qubit set per part) for a .qasm file and its HiSVSIM part file."""
    from quantum_simulations_b200.circuit.qasm import qasm_to_ops
    n, ops, stmt = qasm_to_ops(qasm_text, with_statement_index=True)
    parts = read_part_file(part_text)
    if stmt and max(stmt) >= len(parts):
        raise ValueError(f"the part file labels {len(parts)} gates, the program has {max(stmt) + 1}")
    new_ops, qsets, _ = reorder_by_parts(ops, [parts[s_] for s_ in stmt])
    return n, new_ops, qsets


def qasm_parts(qasm_text: str, part_text: str):
    """(n_qubits, [ops of part 0, ops of part 1, ...]) in execution order — the input of
    circuit.sharding.plan_parts, which runs every part as ONE STAGE with a single gather of the qubits it mixes
    in front of it (HiSVSIM's execution model, hisvsim_repo/execute.hpp:665-685)."""
    from quantum_simulations_b200.circuit.qasm import qasm_to_ops
    n, ops, stmt = qasm_to_ops(qasm_text, with_statement_index=True)
    parts = read_part_file(part_text)
    if stmt and max(stmt) >= len(parts):
        raise ValueError(f"the part file labels {len(parts)} gates, the program has {max(stmt) + 1}")
    labels = [parts[s_] for s_ in stmt]
    new_ops, _, order = reorder_by_parts(ops, labels)
    count = {p: 0 for p in order}
    for lb in labels:
        count[lb] += 1
    out, at = [], 0
    for p in order:
        out.append(new_ops[at:at + count[p]])
        at += count[p]
    return n, out
