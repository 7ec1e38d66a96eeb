"""method="ilp" of ``atlas_stages``: optimal local-qubit sets per stage by integer programming.

Same model as the reference (wenbo_engine/circuit/staging.py:176-315, after Atlas
``compute_local_qubits_with_ilp``), solved with SciPy's HiGHS front end (``scipy.optimize.milp``)
instead of PuLP/CBC:

    x[s][q] in {0,1}   qubit q is local in stage s
    y[g][s] in {0,1}   gate g runs in stage s
    (1) every gate runs in exactly one stage
    (2) a gate never runs before the previous gate on any of its qubits
    (3) the non-insular qubits of a gate are local in its stage
    (4) exactly k qubits are local in every stage
    minimise the number of qubits that change side between consecutive stages

The smallest feasible number of stages is found by bisection below a heuristic upper bound."""
from __future__ import annotations

import numpy as np


def _solve(gates, n: int, k: int, S: int, ni, pred, time_limit: float):
    from scipy.optimize import Bounds, LinearConstraint, milp
    from scipy.sparse import lil_matrix

    G = len(gates)
    nx, ny, nd = S * n, G * S, (S - 1) * n
    X = lambda s, q: s * n + q                      # noqa: E731
    Y = lambda g, s: nx + g * S + s                 # noqa: E731
    D = lambda s, q: nx + ny + s * n + q            # noqa: E731
    nv = nx + ny + nd
    rows, lo, hi = [], [], []

    def add(coefs: dict, lo_, hi_):
        rows.append(coefs)
        lo.append(lo_)
        hi.append(hi_)

    for g in range(G):                                                  # (1)
        add({Y(g, s): 1.0 for s in range(S)}, 1.0, 1.0)
    for g in range(G):                                                  # (2)
        for p in pred[g]:
            for s in range(S):
                c = {Y(p, sp): 1.0 for sp in range(s + 1)}
                c[Y(g, s)] = c.get(Y(g, s), 0.0) - 1.0
                add(c, 0.0, np.inf)
    for g in range(G):                                                  # (3)
        for q in ni[g]:
            for s in range(S):
                add({X(s, q): 1.0, Y(g, s): -1.0}, 0.0, np.inf)
    for s in range(S):                                                  # (4)
        add({X(s, q): 1.0 for q in range(n)}, float(k), float(k))
    for s in range(S - 1):                                              # d >= |x[s] - x[s+1]|
        for q in range(n):
            add({D(s, q): 1.0, X(s, q): -1.0, X(s + 1, q): 1.0}, 0.0, np.inf)
            add({D(s, q): 1.0, X(s, q): 1.0, X(s + 1, q): -1.0}, 0.0, np.inf)
    A = lil_matrix((len(rows), nv))
    for i, c in enumerate(rows):
        for j, v in c.items():
            A[i, j] = v
    cost = np.zeros(nv)
    cost[nx + ny:] = 1.0
    integrality = np.concatenate([np.ones(nx + ny), np.zeros(nd)])
    res = milp(cost, constraints=LinearConstraint(A.tocsr(), np.array(lo), np.array(hi)), integrality=integrality,
               bounds=Bounds(np.zeros(nv), np.concatenate([np.ones(nx + ny), np.full(nd, np.inf)])),
               options={"time_limit": time_limit, "disp": False})
    if res.status != 0 or res.x is None:
        return None
    return [{q for q in range(n) if res.x[X(s, q)] > 0.5} for s in range(S)]


def local_sets_ilp(gates: list[dict], n: int, k: int, max_stages: int | None = None,
                   time_limit: float = 10.0) -> list[set[int]]:
    """Local-qubit set of every stage, fewest stages first, then fewest side changes."""
    from quantum_simulations_b200.circuit.staging import non_insular_qubits

    if not gates:
        return [set(range(min(k, n)))]
    ni = [non_insular_qubits(g) for g in gates]
    if any(len(set(q)) > k for q in ni):
        raise ValueError("a gate needs more local qubits than k")
    last: dict[int, int] = {}
    pred: list[list[int]] = [[] for _ in gates]
    for gi, g in enumerate(gates):
        for q in g["qubits"]:
            if q in last:
                pred[gi].append(last[q])
            last[q] = gi
    hi_s = max_stages if max_stages is not None else len(gates)
    lo_s, best = 1, None
    while lo_s <= hi_s:
        mid = (lo_s + hi_s) // 2
        got = _solve(gates, n, k, mid, ni, pred, time_limit)
        if got is not None:
            best, hi_s = got, mid - 1
        else:
            lo_s = mid + 1
    if best is None:
        raise RuntimeError("staging ILP found no feasible plan")
    return best
