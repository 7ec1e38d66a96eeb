"""Circuit staging with qubit remapping (reference wenbo_engine/circuit/staging.py).

Same public surface — ``non_insular_qubits``, ``QubitMap``, ``atlas_stages``,
``permute_state``, ``staging_stats`` — and the same step IR, but the stage builder here keeps
PER-QUBIT PROGRAM ORDER, which the reference's ``_local_sets_to_steps`` (staging.py:447-519)
does not: it emits all of a stage's local gates before all of its non-local ops, so a
diagonal gate on a global qubit can jump over a later dependent local gate
(11/60 random circuits wrong with method="heuristic").  Here a stage is a
dependency-closed run of gates and a new step is started whenever emitting "local ops, then
non-local ops" would reorder two gates that do not commute.

k = number of local bit positions (log2 of the chunk / shard size).  A gate is executable
in a stage when its non-insular qubits are local; insular (diagonal) use of a global qubit
needs no data movement on a GPU shard (it is a per-rank constant).
"""
from __future__ import annotations

from collections import Counter

import numpy as np

from quantum_simulations_b200.circuit.fusion import batch_levels, fuse_1q_ops
from quantum_simulations_b200.circuit.io import levelize, validate_circuit_dict
from quantum_simulations_b200.kernel import gates as gmod

_SPARSE_GATES = frozenset({"Z", "S", "T", "CZ", "CR"})


def non_insular_qubits(gate: dict) -> list[int]:
    """Qubits of `gate` that MUST be local (reference staging.py:77-98, Atlas is_sparse()):
    none for the diagonal gates Z,S,T,CZ,CR; all of them otherwise."""
    return [] if gate["gate"] in _SPARSE_GATES else list(gate["qubits"])


def mixed_qubits(gate: dict) -> list[int]:
    """Finer than non_insular_qubits: qubits a gate actually MIXES (non-diagonal action).
    R is diagonal too, and controlled gates only mix their target."""
    kind = gmod.gate_structure(gate["gate"])
    if kind == "diag":
        return []
    if kind == "ctrl":
        return [gate["qubits"][1]]
    return list(gate["qubits"])


class QubitMap:
    """Bidirectional logical <-> physical qubit map (reference staging.py:103-131)."""

    def __init__(self, n: int):
        self.n = n
        self._l2p = list(range(n))
        self._p2l = list(range(n))

    def phys(self, logical: int) -> int:
        return self._l2p[logical]

    def logical(self, physical: int) -> int:
        return self._p2l[physical]

    def local_set(self, k: int) -> set[int]:
        return set(self._p2l[: min(k, self.n)])

    def swap_phys(self, pa: int, pb: int) -> None:
        la, lb = self._p2l[pa], self._p2l[pb]
        self._p2l[pa], self._p2l[pb] = lb, la
        self._l2p[la], self._l2p[lb] = pb, pa

    def to_list(self) -> list[int]:
        return list(self._l2p)

    def is_identity(self) -> bool:
        return self._l2p == list(range(self.n))


def _swap_ops(qmap: QubitMap, want_local: set[int], k: int):
    """SWAP ops (on physical positions) that bring `want_local` into positions < k."""
    have = qmap.local_set(k)
    come = sorted(want_local - have)
    go = sorted(have - want_local)
    ops = []
    for lin, lout in zip(come, go):
        pin, pout = qmap.phys(lin), qmap.phys(lout)
        ops.append(([pout, pin], gmod.SWAP()))
        qmap.swap_phys(pout, pin)
    return ops


class _StepBuilder:
    """Accumulates ops in program order into reference-shaped steps without reordering
    non-commuting gates across the local/non-local split of a step."""

    def __init__(self, k: int):
        self.k = k
        self.steps: list[dict] = []
        self._loc: list = []
        self._non: list = []
        self._non_qubits: set[int] = set()

    def add(self, phys_qs: list[int], U: np.ndarray) -> None:
        if all(q < self.k for q in phys_qs):
            # a local op emitted now would execute BEFORE the pending non-local ops of the
            # step: only legal if it shares no qubit with them
            if self._non_qubits & set(phys_qs):
                self.flush()
            self._loc.append((phys_qs, U))
        else:
            self._non.append((phys_qs, U))
            self._non_qubits.update(phys_qs)

    def flush(self) -> None:
        if self._loc or self._non:
            self.steps.append({"local_ops": fuse_1q_ops(self._loc), "nonlocal_ops": list(self._non)})
        self._loc, self._non, self._non_qubits = [], [], set()


def _executable_prefix(gates, done, is_ok):
    """Indices of not-yet-done gates that can run now, in order, respecting dependencies."""
    blocked: set[int] = set()
    out = []
    for i, g in enumerate(gates):
        if done[i]:
            continue
        qs = g["qubits"]
        if not (blocked & set(qs)) and is_ok(g):
            out.append(i)
        else:
            blocked.update(qs)
    return out


def _pick_local_heuristic(gates, done, n, k, local_now: set[int]) -> set[int]:
    """Atlas priority (reference staging.py:320-421): qubits of the first blocked gate, then
    qubits with many pending gates that are not executable, then many pending gates."""
    first, glob_cnt, loc_cnt = set(), Counter(), Counter()
    seen_first = False
    for i, g in enumerate(gates):
        if done[i]:
            continue
        need = non_insular_qubits(g)
        ok = all(q in local_now for q in need)
        for q in g["qubits"]:
            (loc_cnt if ok else glob_cnt)[q] += 1
        if not seen_first:
            first.update(need or g["qubits"])
            seen_first = True
    order = sorted(range(n), key=lambda q: (-(q in first), -glob_cnt[q], -loc_cnt[q], q))
    return set(order[:k])


def _pick_local_greedy(gates, done, n, k, lookahead: int) -> set[int]:
    """Frequency look-ahead (reference staging.py:524-582)."""
    freq: Counter = Counter()
    seen = 0
    for i, g in enumerate(gates):
        if done[i]:
            continue
        for q in g["qubits"]:
            freq[q] += 1
        seen += 1
        if seen >= lookahead:
            break
    want = [q for q, _ in freq.most_common(k)]
    want += [q for q in range(n) if q not in want][: k - len(want)]
    return set(want[:k])


def atlas_stages(circuit_dict: dict, k: int, method: str = "heuristic",
                 lookahead: int = 200) -> tuple[list[dict], list[int]]:
    """Circuit -> (steps, log_to_phys) with SWAP steps between stages
    (reference staging.py:587-634)."""
    cd = validate_circuit_dict(circuit_dict)
    n = cd["number_of_qubits"]
    gates = cd["gates"]
    if n <= k:
        return batch_levels(levelize(cd), k), list(range(n))
    if method not in ("heuristic", "greedy", "ilp"):
        raise ValueError(f"unknown staging method: {method!r}")
    if method == "ilp":
        from quantum_simulations_b200.circuit.staging_ilp import local_sets_ilp
        # upper bound on the number of stages: what the heuristic needs (bisection starts below it)
        heur_steps, _ = atlas_stages(cd, k, method="heuristic", lookahead=lookahead)
        bound = 1 + sum(1 for st in heur_steps if st["nonlocal_ops"] and not st["local_ops"])
        try:
            plan = iter(local_sets_ilp(gates, n, k, max_stages=max(1, bound)))
        except (ValueError, RuntimeError):
            plan = iter(())                              # e.g. a gate wider than k: heuristic picks below
    else:
        plan = None

    qmap = QubitMap(n)
    done = [False] * len(gates)
    steps: list[dict] = []
    guard = 0
    while not all(done):
        guard += 1
        if guard > 4 * len(gates) + 8:
            raise RuntimeError("staging did not converge")
        local_now = qmap.local_set(k)
        run = _executable_prefix(gates, done, lambda g: all(q in local_now for q in non_insular_qubits(g)))
        if not run:
            i_first = next(i for i in range(len(gates)) if not done[i])
            first = gates[i_first]
            need = set(non_insular_qubits(first))
            if len(need) > k:
                # cannot be localised (e.g. a 2-qubit gate with 1 local position): run it as a
                # non-local op, which the reference executes as a chunk-group butterfly
                steps.append({"local_ops": [], "nonlocal_ops": [(
                    [qmap.phys(q) for q in first["qubits"]],
                    gmod.gate_matrix(first["gate"], first["params"]))]})
                done[i_first] = True
                continue
            if plan is not None:
                want = next(plan, None) or _pick_local_heuristic(gates, done, n, k, local_now)
            elif method == "greedy":
                want = _pick_local_greedy(gates, done, n, k, lookahead)
            else:
                want = _pick_local_heuristic(gates, done, n, k, local_now)
            if not need <= want:                       # always make progress on the first gate
                spare = sorted(want - need, reverse=True)
                want = set(need) | set(spare[: k - len(need)])
            swaps = _swap_ops(qmap, want, k)
            if swaps:
                steps.append({"local_ops": [], "nonlocal_ops": swaps})
            continue
        sb = _StepBuilder(k)
        for i in run:
            g = gates[i]
            sb.add([qmap.phys(q) for q in g["qubits"]], gmod.gate_matrix(g["gate"], g["params"]))
            done[i] = True
        sb.flush()
        steps.extend(sb.steps)
    return steps, qmap.to_list()


def permute_state(state: np.ndarray, log_to_phys: list[int]) -> np.ndarray:
    """Physical layout -> logical qubit order (reference staging.py:639-658):
    out[x] = state[y] where bit log_to_phys[q] of y equals bit q of x."""
    n = len(log_to_phys)
    if all(log_to_phys[q] == q for q in range(n)):
        return state
    axes = [0] * n                      # C-order axis j <-> bit n-1-j
    for q, p in enumerate(log_to_phys):
        axes[n - 1 - q] = n - 1 - p
    return np.ascontiguousarray(state.reshape((2,) * n).transpose(axes)).reshape(-1)


def staging_stats(circuit_dict: dict, k: int, method: str = "heuristic") -> dict:
    """Step counts with and without staging (reference staging.py:663-689)."""
    cd = validate_circuit_dict(circuit_dict)
    base = batch_levels(levelize(cd), k)
    staged, _ = atlas_stages(circuit_dict, k, method=method)
    n_base, n_st = len(base), len(staged)
    return {
        "baseline_steps": n_base,
        "staged_steps": n_st,
        "baseline_nonlocal_steps": sum(1 for s in base if s.get("nonlocal_ops")),
        "staged_nonlocal_steps": sum(1 for s in staged if s.get("nonlocal_ops")),
        "reduction": f"{n_base}->{n_st} ({(1 - n_st / max(n_base, 1)) * 100:.0f}% fewer I/O passes)",
    }
