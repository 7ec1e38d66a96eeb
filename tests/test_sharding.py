"""Sharded programs (n_local < n): the stage planner inserts global<->local swaps; executed
here on all shards through the NumPy pass emulator and compared with the oracle."""
import numpy as np
import pytest

from oracle import ref_dense as O
from quantum_simulations_b200.circuit.io import validate_circuit_dict
from quantum_simulations_b200.circuit.passes import PassCompiler, SwapStep
from quantum_simulations_b200.kernel import gates as G
from quantum_simulations_b200 import workloads as W
from tests.pass_emulator import run_program_sharded


def ir_ops(cd):
    cd = validate_circuit_dict(cd)
    return [(g["qubits"], G.gate_matrix(g["gate"], g["params"])) for g in cd["gates"]]


def check(cd, g, **kw):
    n = cd["number_of_qubits"]
    prog = PassCompiler(n, n_local=n - g, **kw).compile(ir_ops(cd))
    psi = np.zeros(1 << n, dtype=np.complex128)
    psi[0] = 1
    psi = run_program_sharded(prog, psi)
    want = O.simulate(validate_circuit_dict(cd))
    if prog.rank_flip_mask:                       # rank_flips=True: shard r holds logical shard r ^ mask
        world = 1 << g
        shards = psi.reshape(world, -1)
        psi = np.concatenate([shards[r ^ prog.rank_flip_mask] for r in range(world)])
    assert prog.final_pos == list(range(n))
    assert not any(f for q, f in enumerate(prog.final_flips) if prog.final_pos[q] < n - g)
    assert np.abs(psi - want).max() <= 1e-12
    return prog


@pytest.mark.parametrize("g", [1, 2, 3])
@pytest.mark.parametrize("workload", ["random_1q_cz", "random_mixed", "qft", "ghz"])
def test_sharded_matches_oracle(workload, g):
    n = 12
    cd = {"random_1q_cz": lambda: W.random_1q_cz(n, 20, 1234), "random_mixed": lambda: W.random_mixed(n, 150, 11),
          "qft": lambda: W.qft(n), "ghz": lambda: W.ghz(n)}[workload]()
    prog = check(cd, g, tile_bits=7, low_bits=2)
    assert prog.stats["swaps"] >= 1


@pytest.mark.parametrize("seed", range(6))
def test_sharded_random_mixed_seeds(seed):
    check(W.random_mixed(11, 120, seed), 2, tile_bits=6, low_bits=2)


@pytest.mark.parametrize("g", [1, 2, 3])
def test_swap_anywhere_names_arbitrary_local_positions(g):
    """The peer-memory swap kernel takes any local positions: no relabel pass before the swap."""
    n = 13
    prog = check(W.random_1q_cz(n, 20, 1234), g, tile_bits=7, low_bits=2, swap_anywhere=True)
    assert prog.stats["swaps"] >= 1
    check(W.random_mixed(n, 150, 3), g, tile_bits=7, low_bits=2, swap_anywhere=True)


@pytest.mark.parametrize("g", [1, 2, 3])
@pytest.mark.parametrize("seed", range(4))
def test_rank_flips_rename_shards_instead_of_moving_data(g, seed):
    n = 12
    check(W.random_1q_cz(n, 20, seed), g, tile_bits=7, low_bits=2, swap_anywhere=True, rank_flips=True)
    check(W.random_mixed(n, 150, seed), g, tile_bits=7, low_bits=2, rank_flips=True)
    gates = [{"qubits": [n - 1], "gate": "X"}, {"qubits": [0], "gate": "H"}, {"qubits": [n - 1, 0], "gate": "CNOT"}]
    prog = check({"number_of_qubits": n, "gates": gates}, 1, tile_bits=7, low_bits=2, rank_flips=True)
    assert prog.rank_flip_mask == 1 and prog.stats["swaps"] == 0


def test_diagonal_only_use_of_rank_bits_needs_no_swap():
    """Z/S/T/CZ/CR on rank-bit qubits are per-shard constants (Atlas 'insular' qubits)."""
    n = 10
    gates = [{"qubits": [q], "gate": "H"} for q in range(n - 2)]
    gates += [{"qubits": [n - 1, 0], "gate": "CZ"}, {"qubits": [n - 2], "gate": "T"},
              {"qubits": [n - 1, 3], "gate": "CR", "params": {"k": 3}}, {"qubits": [n - 2, n - 1], "gate": "CZ"}]
    prog = check({"number_of_qubits": n, "gates": gates}, 2, tile_bits=6, low_bits=2)
    assert prog.stats["swaps"] == 0


def test_x_on_rank_bit_is_materialised():
    n = 9
    gates = [{"qubits": [n - 1], "gate": "X"}, {"qubits": [0], "gate": "H"}, {"qubits": [n - 1, 0], "gate": "CNOT"}]
    check({"number_of_qubits": n, "gates": gates}, 1, tile_bits=6, low_bits=2)


@pytest.mark.parametrize("g", [1, 2, 3])
@pytest.mark.parametrize("workload", ["random_1q_cz", "random_mixed", "qft"])
def test_planned_placements_incl_exchange_candidates_match_oracle(workload, g):
    """sharding.plan from |0...0>: the initial placement is chosen among block and (with swap_anywhere)
    plain-exchange candidates; whatever wins must give the oracle's state in the identity layout."""
    from quantum_simulations_b200.circuit import sharding
    picked_exchange = 0
    for n, seed in ((14, 1234), (15, 5), (16, 77)):
        cd = {"random_1q_cz": lambda: W.random_1q_cz(n, 20, seed), "random_mixed": lambda: W.random_mixed(n, 160, seed),
              "qft": lambda: W.qft(n)}[workload]()
        kw = dict(tile_bits=7, low_bits=2, swap_anywhere=True, rank_flips=True)
        prog = sharding.plan(ir_ops(cd), n, n - g, **kw)
        base = sharding.candidate_placements(n, g)
        assert prog.stats["init_pos"] in sharding.candidate_placements(n, g, direct=True)
        picked_exchange += prog.stats["init_pos"] not in base
        psi = np.zeros(1 << n, dtype=np.complex128)
        psi[0] = 1
        psi = run_program_sharded(prog, psi)
        if prog.rank_flip_mask:
            shards = psi.reshape(1 << g, -1)
            psi = np.concatenate([shards[r ^ prog.rank_flip_mask] for r in range(1 << g)])
        assert prog.final_pos == list(range(n))
        assert np.abs(psi - O.simulate(validate_circuit_dict(cd))).max() <= 1e-12
    # every exchange candidate on its own, not only when the cost model picks it
    n = 14
    cd = {"random_1q_cz": lambda: W.random_1q_cz(n, 20, 3), "random_mixed": lambda: W.random_mixed(n, 160, 3),
          "qft": lambda: W.qft(n)}[workload]()
    extra = [p for p in sharding.candidate_placements(n, g, direct=True) if p not in sharding.candidate_placements(n, g)]
    assert extra
    for init in extra:
        prog = PassCompiler(n, n - g, tile_bits=7, low_bits=2, swap_anywhere=True, rank_flips=True).compile(
            ir_ops(cd), init_pos=init, home_pos=list(range(n)))
        psi = np.zeros(1 << n, dtype=np.complex128)
        psi[0] = 1
        psi = run_program_sharded(prog, psi)
        if prog.rank_flip_mask:
            shards = psi.reshape(1 << g, -1)
            psi = np.concatenate([shards[r ^ prog.rank_flip_mask] for r in range(1 << g)])
        assert np.abs(psi - O.simulate(validate_circuit_dict(cd))).max() <= 1e-12


@pytest.mark.parametrize("method", ["heuristic", "greedy", "ilp"])
def test_atlas_stages_drive_the_sharded_plan(method):
    """sharding.plan_atlas: the reference's staging (atlas_stages: local sets per stage, SWAP steps, log_to_phys)
    supplies the stage boundaries of the sharded program; the state comes out in atlas's physical layout and
    permute_state maps it back (reference staging.py:587-658)."""
    from quantum_simulations_b200.circuit import sharding
    from quantum_simulations_b200.circuit.staging import permute_state
    for n, g, cd in ((10, 2, W.random_mixed(10, 120, 3)), (11, 1, W.qft(11)), (12, 3, W.random_1q_cz(12, 12, 7))):
        cd = validate_circuit_dict(cd)
        prog, l2p = sharding.plan_atlas(cd, n - g, method=method, tile_bits=6, low_bits=2, swap_anywhere=True)
        assert sorted(l2p) == list(range(n)) and prog.stats["parts"] >= 1
        psi = np.random.default_rng(1).standard_normal(1 << n) + 0j      # the fused first pass ignores what is there
        if not prog.fused_init:
            psi = np.zeros(1 << n, dtype=np.complex128)
            psi[0] = 1
        got = permute_state(run_program_sharded(prog, psi), l2p)
        assert np.abs(got - O.simulate(cd)).max() <= 1e-12
