"""Race check of the generated pass kernels on the CPU: tests/jit_tsan.py builds the kernel + host prelude with
g++ -fsanitize=thread and runs it with one OS thread per CUDA thread.  Clean kernels of every load form (single, paired,
four tiles at once), both dtypes, the zero-input form, chunk-sized launches with tile blocks, table phases, the
warp-local round exchange and the 4-group / 7-buffer variants must produce NO report; three kernels whose
synchronisation is broken on purpose must be reported (the check can see what it is looking for)."""
import re

import numpy as np
import pytest

from quantum_simulations_b200 import workloads as W
from quantum_simulations_b200.kernel.cuda_dense import compile_circuit
from tests.jit_tsan import race_check, tsan_available

pytestmark = pytest.mark.skipif(not tsan_available(), reason="g++ -fsanitize=thread does not build / run here")

N = 16                      # 32 tiles per launch: every ring buffer is refilled several times


@pytest.fixture(scope="module")
def prog():
    return compile_circuit(W.random_1q_cz(N, 20, 1234), zero_init=True)


def _clean(step, **kw):
    count, text = race_check(step, N, **kw)
    assert count == 0, text[-3000:]


def test_default_kernels_of_a_planned_run_are_race_free(prog):
    assert prog.fused_init
    for step in prog.passes[:4]:
        if step.desc.zero_input:                       # the zero-fill form launches the kernel on the one live tile
            _clean(step, tile_range=(0, 1))
            _clean(step)                               # and QSV_INIT_PASS_FULL=1 launches it over every tile
        else:
            _clean(step, grid=2)


@pytest.mark.parametrize("pair,blk,grid", [("0", 0, 3), ("1", 1, 2), ("1", 3, 1), ("2", 2, 2)])
def test_load_forms_and_tile_blocks(monkeypatch, prog, pair, blk, grid):
    monkeypatch.setenv("QSV_JIT_PAIR", pair)
    _clean(prog.passes[1], grid=grid, tile_block=blk)
    _clean(prog.passes[1], grid=grid, tile_block=blk, tile_range=(3, 14))        # a chunk-sized launch with an odd tile count


def test_complex64_and_table_phases():
    q = compile_circuit(W.qft(N), dtype="complex64", zero_init=False)
    _clean(q.passes[0], dtype="complex64", grid=2)
    q = compile_circuit(W.qft(N), zero_init=False)
    assert q.passes[0].tables is not None
    _clean(q.passes[0], grid=2)


def test_generator_variants(monkeypatch):
    monkeypatch.setenv("QSV_JIT_WARP_SYNC", "1")
    p = compile_circuit(W.random_1q_cz(N, 20, 1234), zero_init=False, warp_local_rounds=True)
    from tests.jit_host_run import kernel_source
    step = next(s for s in p.passes if "__syncwarp();" in kernel_source(s))
    _clean(step, grid=2)
    monkeypatch.delenv("QSV_JIT_WARP_SYNC")
    monkeypatch.setenv("QSV_JIT_GROUPS", "4")
    _clean(p.passes[0], grid=1)
    monkeypatch.delenv("QSV_JIT_GROUPS")
    monkeypatch.setenv("QSV_JIT_NBUF", "7")
    _clean(p.passes[0], grid=1)


@pytest.mark.parametrize("what", ["no group barrier between rounds", "consumer does not wait for the fill",
                                  "buffer handed back before it is read", "every round exchange behind a warp barrier"])
def test_broken_synchronisation_is_reported(prog, what):
    """positive controls: the same kernel with ONE piece of its synchronisation removed"""
    step = prog.passes[1]
    assert step.desc.n_rounds >= 2

    def early_release(src):
        src = src.replace("mbar_arrive(&S.empty[b]);", "")
        assert "mbar_wait(&S.full[b], use & 1);" in src
        return src.replace("mbar_wait(&S.full[b], use & 1);", "mbar_wait(&S.full[b], use & 1); mbar_arrive(&S.empty[b]);")

    mutate = {"no group barrier between rounds": lambda s: s.replace("group_bar(grp);", ""),
              "consumer does not wait for the fill": lambda s: s.replace("mbar_wait(&S.full[b], use & 1);", ""),
              "buffer handed back before it is read": early_release,
              "every round exchange behind a warp barrier": lambda s: s.replace("group_bar(grp);", "__syncwarp();")}[what]
    count, text = race_check(step, N, mutate=mutate, timeout=300, stop_at_first=True)
    assert count > 0, what
    assert "data race" in text


def test_tiles_are_closed_address_sets():
    """What the race check above cannot see (the host launcher runs the CTAs of a grid one after the other): two CTAs
    never touch the same address, because a pass stores a tile into exactly the address set it loaded it from — the
    store positions of every pass are a permutation of its load positions, so the tiles of a launch are disjoint."""
    from quantum_simulations_b200.circuit import sharding
    from quantum_simulations_b200.circuit.io import validate_circuit_dict
    from quantum_simulations_b200.kernel.cuda_dense import circuit_ops
    progs = [compile_circuit(W.random_1q_cz(20, 20, 1234)), compile_circuit(W.qft(18)), compile_circuit(W.random_mixed(15, 300, 3)),
             sharding.plan(circuit_ops(validate_circuit_dict(W.random_1q_cz(22, 20, 1234))), 22, 19, swap_anywhere=True, rank_flips=True),
             sharding.plan(circuit_ops(validate_circuit_dict(W.random_1q_cz(30, 20, 1234))), 30, 29, search=24, swap_anywhere=True, rank_flips=True)]
    seen = 0
    for p in progs:
        for s in p.passes:
            t = s.desc.n_tile
            load, store = list(s.desc.load_bits[:t]), list(s.desc.store_bits[:t])
            assert sorted(load) == sorted(store) and len(set(load)) == t
            inside = sum(1 << b for b in store)
            assert s.desc.store_flip & ~inside == 0                # a store flip only permutes addresses INSIDE the tile
            seen += 1
    assert seen > 25


def test_kernels_stay_inside_the_shard_and_the_ring(monkeypatch, prog):
    """The same host build under AddressSanitizer + UBSan (the host-model counterpart of compute-sanitizer memcheck): the
    state is ONE heap block of exactly 2^n amplitudes, so a global index outside the shard is reported — none for the
    default kernels, paired loads with tile blocks and an odd chunk-sized launch, complex64; a launch that claims more tiles
    than the shard holds is the positive control."""
    san = "address,undefined"
    for step in prog.passes[:3]:
        count, text = race_check(step, N, grid=2, sanitizer=san, tile_range=(0, 1) if step.desc.zero_input else None)
        assert count == 0, text[-3000:]
    count, text = race_check(prog.passes[1], N, grid=3, tile_block=2, tile_range=(5, 18), sanitizer=san)
    assert count == 0, text[-3000:]
    q = compile_circuit(W.qft(N), dtype="complex64", zero_init=False)
    count, text = race_check(q.passes[0], N, dtype="complex64", grid=2, sanitizer=san)
    assert count == 0, text[-3000:]
    count, text = race_check(prog.passes[1], N, grid=2, tile_range=(0, 40), sanitizer=san)     # 32 tiles exist
    assert count > 0 and "heap-buffer-overflow" in text
