"""Seeded fuzz of the pass compiler / stage planner against the oracle through the NumPy pass
emulator: random sizes, tile shapes, shard counts and every planner switch (absorbed pre-ops, table
phases, swaps at arbitrary positions, rank-bit X frames).  20,000 seeds of this generator ran clean
when it was written; a slice runs in the suite."""
import numpy as np
import pytest

from oracle import ref_dense as O
from quantum_simulations_b200 import workloads as W
from quantum_simulations_b200.circuit import sharding
from quantum_simulations_b200.circuit.io import validate_circuit_dict
from quantum_simulations_b200.circuit.passes import PassCompiler
from quantum_simulations_b200.kernel import gates as G
from tests.pass_emulator import run_program, run_program_sharded


def _case(seed: int):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(6, 13))
    t = int(rng.integers(5, min(n, 9) + 1))
    a = int(rng.integers(0, min(4, t - 4) + 1))
    g = int(rng.integers(0, min(3, n - t) + 1)) if n - t > 0 else 0
    kind = int(rng.integers(0, 3))
    cd = [W.random_mixed(n, int(rng.integers(40, 200)), seed), W.random_1q_cz(n, int(rng.integers(4, 16)), seed),
          W.qft(n)][kind]
    kw = dict(tile_bits=t, low_bits=a, max_rounds=int(rng.integers(2, 6)), swap_anywhere=bool(rng.integers(0, 2)),
              rank_flips=bool(rng.integers(0, 2)), table_phases=bool(rng.integers(0, 2)), absorb=bool(rng.integers(0, 2)))
    planned = bool(rng.integers(0, 2))
    kw["low_store_round"] = bool(rng.integers(0, 2))        # drawn last: earlier seeds keep their cases
    kw["warp_local_rounds"] = bool(rng.integers(0, 2))
    return n, g, validate_circuit_dict(cd), kw, planned


@pytest.mark.parametrize("block", range(8))
def test_fuzz_block(block):
    for seed in list(range(block * 40, block * 40 + 40)) + [5746, 7444]:      # + two seeds that once failed
        n, g, cd, kw, planned = _case(seed)
        ops = [(q["qubits"], G.gate_matrix(q["gate"], q["params"])) for q in cd["gates"]]
        # planned: free initial placement from |0...0> (sharding.plan / plan_single), every other seed
        # with zero-support skipping on one device
        if planned and g == 0:
            prog = sharding.plan_single(ops, n, "complex128", True, seed % 2 == 0, **kw)
        elif planned:
            prog = sharding.plan(ops, n, n - g, **kw)
        else:
            prog = PassCompiler(n, n - g, **kw).compile(ops)
        psi = np.zeros(1 << n, dtype=np.complex128)
        psi[0] = 1
        if prog.fused_init:                 # the first pass creates |0...0> itself: whatever was there is ignored
            psi = np.random.default_rng(seed).standard_normal(1 << n) + 0j
        psi = run_program_sharded(prog, psi) if g else run_program(prog, psi)
        if prog.rank_flip_mask:
            shards = psi.reshape(1 << g, -1)
            psi = np.concatenate([shards[r ^ prog.rank_flip_mask] for r in range(1 << g)])
        assert prog.final_pos == list(range(n)), (seed, kw)
        assert np.abs(psi - O.simulate(cd)).max() <= 1e-12, (seed, kw)
