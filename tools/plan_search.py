"""Explore the seeded variations of the stage planner for one sharded configuration (CPU only): how many passes the
plans of each exploration setting need, and which of the better ones sharding.plan's search would accept.
    python tools/plan_search.py [n_qubits] [log2 shards] [seeds per setting]
e.g. `python tools/plan_search.py 34 1 150`: the default exploration (p = 0.4, first 3 candidates) has 8-pass plans only
with a second swap; with the first 5 candidates three acceptable 8-pass plans exist (sharding.SEARCH_STAGES)."""
import collections
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from quantum_simulations_b200 import workloads as W                                                     # noqa: E402
from quantum_simulations_b200.circuit import sharding                                                  # noqa: E402
from quantum_simulations_b200.circuit.io import validate_circuit_dict                                  # noqa: E402
from quantum_simulations_b200.circuit.passes import PassCompiler                                       # noqa: E402
from quantum_simulations_b200.circuit.sharding import (_plan_key, candidate_placements, estimate_seconds_v2,   # noqa: E402
                                                       plan_transitions)
from quantum_simulations_b200.kernel.cuda_dense import circuit_ops                                     # noqa: E402

SETTINGS = ((0.4, 3), (0.25, 3), (0.6, 3), (0.4, 5), (0.7, 4), (0.15, 2))


def main() -> int:
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 34
    g = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    seeds = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    ops = circuit_ops(validate_circuit_dict(W.random_1q_cz(n, 20, 1234)))
    kw = dict(swap_anywhere=True, rank_flips=True, max_rounds=3)
    base = PassCompiler(n, n - g, "complex128", **kw)
    greedy = sharding.plan(ops, n, n - g, search=0, **kw)
    init0 = greedy.stats["init_pos"]
    ident = list(range(n))
    ref = base.compile(ops, init_pos=init0, home_pos=ident)          # (not fused: compared like the search compares)
    mcp = 8 if n - g >= 24 else 5
    tr0 = plan_transitions(ref, min_chunk_pos=mcp)
    k0, t0 = _plan_key(ref, tr0), estimate_seconds_v2(ref, tr0)
    print(f"greedy plan: key {k0}, estimate {t0 * 1e3:.1f} ms")
    places = [init0] + [p for p in candidate_placements(n, g, direct=True) if p != init0]
    better = []
    t_start = time.time()
    for p_, k_ in SETTINGS:
        hist = collections.Counter()
        for seed in range(seeds):
            for pi, init in enumerate(places):
                c = PassCompiler(n, n - g, "complex128", **dict(kw, explore_seed=seed, explore_p=p_, explore_k=k_))
                c._lowered = base._lowered
                try:
                    prog = c.compile(ops, init_pos=init, home_pos=ident)
                except (NotImplementedError, RuntimeError):
                    continue
                hist[prog.stats["passes"]] += 1
                if prog.stats["passes"] < k0[0]:
                    tr = plan_transitions(prog, min_chunk_pos=mcp)
                    k = _plan_key(prog, tr)
                    ok = not (k[1] > k0[1] or k[2] > k0[2] or k[3] > k0[3] or k[4] < min(k0[4], 2 * (k[1] - k[3]))
                              or (k0[5] and k[5] < min(k0[5], 3)))
                    t = estimate_seconds_v2(prog, tr)
                    better.append((prog.stats["passes"], round(t * 1e3, 1), ok and t <= 0.96 * t0, p_, k_, seed, pi, k))
        print(f"p = {p_}, first {k_} candidates: passes -> plans {dict(sorted(hist.items()))}   ({time.time() - t_start:.0f} s)", flush=True)
    better.sort()
    print("plans with fewer passes (passes, estimate ms, acceptable, p, k, seed, placement, key):")
    for b in better[:20]:
        print("  ", b)
    print(f"{len(better)} with fewer passes, {sum(b[2] for b in better)} acceptable to the search")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
