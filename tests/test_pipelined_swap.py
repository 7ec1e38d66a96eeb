"""Sharded execution on ONE GPU: every shard of a 2/4/8-way sharded state is a handle of one process on
device 0 (tests/one_process_shards.py).  Covers the exchange kernels of csrc/xchg.cuh (TMA bulk copies
and flag-word ordering), plain swaps and pipelined stage transitions (qsv_swap_pipelined) against the
oracle — the multi-GPU data path, visible to a single-GPU test box.  The reference's own trick for the
same purpose is tiny chunks that force every non-local path (wenbo_engine/tests/test_nonlocal.py:24-50)."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _run(cases, timeout=180):
    env = dict(os.environ, CUDA_DEVICE_MAX_CONNECTIONS="32")
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "one_process_shards.py"), json.dumps(cases)],
                       capture_output=True, text=True, timeout=timeout, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines and lines[-1].get("ok")
    return lines[:-1]


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_pipelined_transitions_match_oracle_c128(world):
    n = 19 + world.bit_length() - 1
    out = _run([{"n": n, "world": world, "dtype": "complex128", "circuit": c} for c in ("random", "qft", "mixed", "ghz")])
    assert any(l["pipeline"] and l["pipelined"] >= 1 for l in out)
    assert all(l["max_abs_err"] <= 1e-12 for l in out)


@pytest.mark.gpu
def test_pipelined_transitions_match_oracle_c64():
    out = _run([{"n": 20, "world": 2, "dtype": "complex64", "circuit": c} for c in ("random", "qft")])
    assert all(l["max_abs_err"] <= 2e-5 for l in out)


@pytest.mark.gpu
def test_parts_as_stages_on_the_device():
    """sharding.plan_parts (HiSVSIM: one stage per part, one gather in front of it) executed on 4 shards."""
    out = _run([{"n": 20, "world": 4, "dtype": "complex128", "circuit": c, "parts": True, "expect_pipelined": False}
                for c in ("random", "mixed")])
    assert all(l["max_abs_err"] <= 1e-12 and l["swaps"] >= 2 for l in out)


def test_transition_plans_are_valid():
    """CPU: every planned transition names chunk bits outside the tiles of its passes and the swapped bits,
    never shares a pass between two swaps, and skips the fused-initialisation pass."""
    from quantum_simulations_b200 import workloads as W
    from quantum_simulations_b200.circuit import sharding
    from quantum_simulations_b200.circuit.io import validate_circuit_dict
    from quantum_simulations_b200.circuit.passes import PassStep, SwapStep
    from quantum_simulations_b200.kernel.cuda_dense import circuit_ops

    seen_any = False
    for n, g, cd in [(31, 1, W.random_1q_cz(31, 20, 1234)), (33, 3, W.random_1q_cz(33, 20, 1234)),
                     (36, 3, W.random_1q_cz(36, 20, 1234)), (34, 2, W.random_1q_cz(34, 20, 1234)),
                     (20, 2, W.qft(20)), (21, 3, W.random_mixed(21, 300, 8)), (20, 1, W.ghz(20))]:
        prog = sharding.plan(circuit_ops(validate_circuit_dict(cd)), n, n - g, "complex128", swap_anywhere=True, rank_flips=True)
        plans = sharding.plan_transitions(prog, min_chunk_pos=8 if n - g >= 24 else 5)
        used = set()
        for k, tr in plans.items():
            seen_any = True
            sw = prog.steps[k]
            assert isinstance(sw, SwapStep) and tr.a_count + tr.b_count >= 1 and 1 <= len(tr.chunk_bits) <= 4
            assert tr.chunk_bits == sorted(set(tr.chunk_bits)) and not set(tr.chunk_bits) & set(sw.local_bits)
            idx = [k - 1 - i for i in range(tr.a_count)] + [k + 1 + i for i in range(tr.b_count)]
            for i in idx:
                st = prog.steps[i]
                assert isinstance(st, PassStep) and i not in used and not st.desc.zero_input and st.desc.n_active < 0
                assert not set(tr.chunk_bits) & set(st.desc.load_bits[: st.desc.n_tile])
                used.add(i)
            assert all(0 <= b < n - g for b in tr.chunk_bits)
    assert seen_any
