"""Workload spec strings shared by oracle/make_golden.py and the tests.

"ghz:10" -> workloads.ghz(10); "random_1q_cz:n:depth:seed"; "random_mixed:n:gates:seed".
"""
from __future__ import annotations

from quantum_simulations_b200 import workloads as W

_GENERATORS = {
    "bell_2q": W.bell_2q, "x_on_q0_3q": W.x_on_q0_3q, "ry_theta": W.ry_theta,
    "cr3_encoded": W.cr3_encoded, "ghz": W.ghz, "qft": W.qft,
    "hadamard_wall": W.hadamard_wall, "random_1q_cz": W.random_1q_cz,
    "random_mixed": W.random_mixed,
}


def circuit_from_spec(spec: str) -> dict:
    name, *args = spec.split(":")
    return _GENERATORS[name](*(int(a) for a in args))


# states frozen from wenbo_engine.kernel.ref_dense.simulate
STATE_SPECS = (
    ["bell_2q", "x_on_q0_3q", "ry_theta", "cr3_encoded", "hadamard_wall:4"]
    + [f"ghz:{n}" for n in (3, 4, 6, 10)]
    + [f"qft:{n}" for n in (2, 3, 4, 5, 8, 11)]
    + [f"random_1q_cz:{n}:20:1234" for n in (4, 7, 10, 12)]
    + [f"random_mixed:{n}:{g}:{s}" for n, g, s in
       ((2, 30, 1), (3, 60, 2), (5, 120, 3), (6, 150, 4), (8, 200, 5), (9, 240, 6),
        (10, 260, 7), (11, 300, 8))]
)

# (spec, chunk_size, reference-runner kwargs) frozen from wenbo_engine.runner.single_node.run
RUNNER_SPECS = (
    ("ghz:4", 4, {}),
    ("qft:4", 2, {}),
    ("qft:5", 8, {"use_fusion": True}),
    ("random_mixed:6:80:21", 8, {}),
    ("random_mixed:6:80:21", 16, {"use_fusion": True}),
    ("random_mixed:5:60:22", 4, {"use_staging": True, "staging_method": "greedy"}),
    ("random_1q_cz:6:8:3", 8, {"use_staging": True, "staging_method": "greedy"}),
)
