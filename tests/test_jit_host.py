"""The run-time specialised kernels, executed on the CPU (tests/jit_host_run.py: the generated CUDA
source compiled with g++ against a host stand-in for the device prelude, one OS thread per CUDA
thread).  Compared with the NumPy pass emulator, which the GPU kernels are compared with on hardware."""
import numpy as np
import pytest

from quantum_simulations_b200 import workloads as W
from quantum_simulations_b200.circuit import sharding
from quantum_simulations_b200.circuit.io import validate_circuit_dict
from quantum_simulations_b200.circuit.passes import PassCompiler, PassStep, SwapStep
from quantum_simulations_b200.kernel import gates as G
from quantum_simulations_b200.kernel.cuda_dense import compile_circuit
from tests.jit_host_run import run_pass_on_host
from tests.pass_emulator import run_pass, swap_bits_full


def _random_state(n, seed, dtype=np.complex128):
    rng = np.random.default_rng(seed)
    psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    return (psi / np.linalg.norm(psi)).astype(dtype)


def _ops(cd):
    cd = validate_circuit_dict(cd)
    return [(g["qubits"], G.gate_matrix(g["gate"], g["params"])) for g in cd["gates"]]


@pytest.mark.parametrize("workload", ["random_1q_cz", "random_mixed", "qft"])
def test_specialised_kernels_match_the_emulator(workload):
    n = 13
    cd = {"random_1q_cz": lambda: W.random_1q_cz(n, 20, 1234), "random_mixed": lambda: W.random_mixed(n, 220, 5),
          "qft": lambda: W.qft(n)}[workload]()
    prog = compile_circuit(cd, zero_init=False)
    psi = _random_state(n, 1)
    for step in prog.passes[:4]:
        want, got = psi.copy(), psi.copy()
        run_pass(want, step.desc, step.ops, n, 0, step.tables)
        run_pass_on_host(step, got, n, grid=2)
        assert np.abs(got - want).max() <= 1e-13
        psi = want


def test_complex64_kernel_and_low_position_register_stores():
    n = 13
    prog = compile_circuit(W.random_1q_cz(n, 20, 7), dtype="complex64", zero_init=False, low_store_round=False)
    psi = _random_state(n, 2)
    for step in prog.passes[:2]:
        want = psi.copy()
        run_pass(want, step.desc, step.ops, n, 0, step.tables)
        got = psi.astype(np.complex64)
        run_pass_on_host(step, got, n, grid=1)
        assert np.abs(got - want).max() <= 2e-6
        psi = want
    prog = compile_circuit(W.random_1q_cz(n, 20, 7), zero_init=False, low_store_round=False)
    assert prog.stats["rounds"] < compile_circuit(W.random_1q_cz(n, 20, 7), zero_init=False, low_store_bits=None).stats["rounds"]
    psi = _random_state(n, 3)
    for step in prog.passes[:3]:
        want, got = psi.copy(), psi.copy()
        run_pass(want, step.desc, step.ops, n, 0, step.tables)
        run_pass_on_host(step, got, n, grid=2)
        assert np.abs(got - want).max() <= 1e-13
        psi = want


def test_zero_input_pass_full_launch_and_zero_fill_plus_one_tile():
    """The first pass of a run from |0...0> ignores what the shard holds.  Two launch forms must give the
    same shard: every tile (the round-1 form) and zero-fill + the one tile that holds amplitude 0."""
    n = 14
    prog = compile_circuit(W.random_1q_cz(n, 20, 1234))
    assert prog.fused_init and prog.passes[0].desc.zero_input == 1
    step = prog.passes[0]
    want = np.zeros(1 << n, dtype=np.complex128)
    want[0] = 1
    run_pass(want, step.desc, step.ops, n, 0, step.tables)
    full = _random_state(n, 4)                       # junk that must be ignored
    run_pass_on_host(step, full, n, grid=3)
    assert np.abs(full - want).max() <= 1e-13
    one = np.zeros(1 << n, dtype=np.complex128)      # = cudaMemsetAsync
    run_pass_on_host(step, one, n, grid=1, tile_range=(0, 1))
    assert np.abs(one - want).max() <= 1e-13
    other_rank = _random_state(n, 5)                 # a shard of rank != 0 is all zero after the pass
    run_pass_on_host(step, other_rank, n, rank=1, grid=2)
    assert not other_rank.any()


@pytest.mark.parametrize("g", [1, 2])
def test_scatter_kernels_do_the_pass_and_the_swap(g):
    """qsv_pass_scatter on the CPU: every rank runs the scatter variant of the pass before a swap; its
    stores go to the second buffers of all ranks.  Afterwards the second buffers must hold what
    'pass, then swap' leaves in the shards."""
    n = 13 + g
    n_loc = n - g
    world = 1 << g
    checked = 0
    for cd in (W.random_1q_cz(n, 20, 1234), W.qft(n), W.random_mixed(n, 200, 9)):
        prog = sharding.plan(_ops(cd), n, n_loc, swap_anywhere=True, rank_flips=True)
        state = np.zeros(1 << n, dtype=np.complex128)
        state[0] = 1
        steps = list(prog.steps)
        for k, step in enumerate(steps):
            shards = state.reshape(world, -1)
            if isinstance(step, SwapStep):
                state = swap_bits_full(state, n, step.global_bits, step.local_bits)
                continue
            nxt = steps[k + 1] if k + 1 < len(steps) else None
            if isinstance(nxt, SwapStep) and checked < 2:
                gb, lb = list(nxt.global_bits), list(nxt.local_bits)
                second = [np.full(1 << n_loc, np.nan + 0j) for _ in range(world)]      # every slot must be written
                for r in range(world):
                    targets, keep = [], 0
                    for i, (gbit, lbit) in enumerate(zip(gb, lb)):
                        keep |= ((r >> (gbit - n_loc)) & 1) << lbit
                    for x in range(1 << len(gb)):
                        rr = r
                        for i, gbit in enumerate(gb):
                            rb = gbit - n_loc
                            rr = (rr & ~(1 << rb)) | (((x >> i) & 1) << rb)
                        targets.append(second[rr])
                    run_pass_on_host(step, shards[r].copy(), n_loc, rank=r, grid=2, scatter=(lb, targets, keep))
                want = state.copy().reshape(world, -1)
                for r in range(world):
                    run_pass(want[r], step.desc, step.ops, n_loc, r, step.tables)
                want = swap_bits_full(want.reshape(-1), n, gb, lb)
                got = np.concatenate(second)
                assert not np.isnan(got).any()
                assert np.abs(got - want).max() <= 1e-13
                checked += 1
            for r in range(world):
                run_pass(shards[r], step.desc, step.ops, n_loc, r, step.tables)
    assert checked == 2


def test_whole_program_through_host_kernels_equals_the_oracle():
    """Every pass of a planned run from |0...0> (fused initialisation in its zero-fill form, table phases,
    pre-ops, store flips, free initial placement) on the host-built specialised kernels."""
    from oracle import ref_dense as O
    n = 14
    cd = validate_circuit_dict(W.random_mixed(n, 260, 21))
    prog = compile_circuit(cd)
    assert prog.fused_init and all(isinstance(s, PassStep) for s in prog.steps)
    psi = _random_state(n, 6)
    for i, step in enumerate(prog.passes):
        if step.desc.zero_input:
            psi[:] = 0
            run_pass_on_host(step, psi, n, grid=1, tile_range=(0, 1))
        else:
            run_pass_on_host(step, psi, n, grid=2)
    assert np.abs(psi - O.simulate(cd)).max() <= 1e-12


def test_warp_local_rounds_with_syncwarp(monkeypatch):
    """PassCompiler(warp_local_rounds=True) keeps the tile positions behind thread bits 5, 6 (the warp
    inside a consumer group) fixed across a round boundary where it can; with QSV_JIT_WARP_SYNC=1 the
    kernel then replaces the group barrier of that boundary by __syncwarp().  On the host every warp is
    32 free-running OS threads, so a boundary that is wrongly declared warp-local gives wrong amplitudes
    (checked once by replacing EVERY barrier: errors of 0.03-0.1)."""
    from tests.jit_host_run import kernel_source
    monkeypatch.setenv("QSV_JIT_WARP_SYNC", "1")
    n = 13
    prog = compile_circuit(W.random_1q_cz(n, 20, 1234), zero_init=False, warp_local_rounds=True)
    assert prog.stats["rounds"] == compile_circuit(W.random_1q_cz(n, 20, 1234), zero_init=False).stats["rounds"]
    psi = _random_state(n, 8)
    local = 0
    for step in prog.passes[:3]:
        local += kernel_source(step).count("__syncwarp();")
        want, got = psi.copy(), psi.copy()
        run_pass(want, step.desc, step.ops, n, 0, step.tables)
        run_pass_on_host(step, got, n, grid=2)
        assert np.abs(got - want).max() <= 1e-13
        psi = want
    assert local >= 2
    monkeypatch.delenv("QSV_JIT_WARP_SYNC")
    assert "__syncwarp();" not in kernel_source(prog.passes[0])          # the default kernels are unchanged


@pytest.mark.parametrize("pair,blk", [("0", 0), ("1", 1), ("2", 2)])
def test_a_pass_launched_chunk_by_chunk_equals_the_whole_pass(monkeypatch, pair, blk):
    """Pipelined stage transitions launch a pass once per CHUNK: index bits outside the tile are fixed through
    the kernel's JitFix argument (positions in tile-index space) and the tile counter runs over the rest.
    The union of the chunk launches must be the pass."""
    n = 14
    monkeypatch.setenv("QSV_JIT_PAIR", pair)
    prog = compile_circuit(W.random_1q_cz(n, 20, 1234), zero_init=False)
    psi = _random_state(n, 5)
    step = prog.passes[1]
    tile = set(step.desc.load_bits[: step.desc.n_tile])
    outside = [p for p in range(n) if p not in tile]
    assert len(outside) == 3
    for chunk_bits in ([outside[0]], [outside[2]], [outside[0], outside[2]], outside):
        want, got = psi.copy(), psi.copy()
        run_pass(want, step.desc, step.ops, n, 0, step.tables)
        pos = [b - sum(t < b for t in tile) for b in chunk_bits]
        tiles = (1 << (n - 11)) >> len(chunk_bits)
        for j in range(1 << len(chunk_bits)):
            val = sum(((j >> i) & 1) << pos[i] for i in range(len(chunk_bits)))
            before = got.copy()
            run_pass_on_host(step, got, n, grid=1 + blk, tile_range=(0, tiles), fix=(pos, val), tile_block=blk)
            changed = np.nonzero(got != before)[0]
            for i, b in enumerate(chunk_bits):                  # a chunk launch stays inside its chunk
                assert np.all((changed >> b) & 1 == (j >> i) & 1)
        assert np.abs(got - want).max() <= 1e-13


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_tile_blocks_and_paired_loads(monkeypatch, dtype):
    """QSV_JIT_TILE_BLOCK (a CTA is dealt 2^k consecutive tiles at a time: JitFix.blk / jit_seq) and QSV_JIT_PAIR=1
    (the producers fill two ring buffers at once) change which CTA moves which tile and how the tile reaches shared
    memory — never the result.  Grids that do not divide the tile count, a single tile, an odd tile range and a
    chunked launch (JitFix positions) are all walked."""
    n = 15                                                  # 16 tiles
    tol = 1e-13 if dtype == "complex128" else 2e-6
    prog = compile_circuit(W.random_1q_cz(n, 20, 1234), dtype=dtype, zero_init=False)
    steps = prog.passes[:2]
    psi = _random_state(n, 6)
    for pair in ("0", "1", "2"):
        monkeypatch.setenv("QSV_JIT_PAIR", pair)
        for step in steps:
            want = psi.copy()
            run_pass(want, step.desc, step.ops, n, 0, step.tables)
            for grid, blk in ((1, 0), (3, 0), (3, 1), (2, 2), (5, 1), (3, 4)):
                got = psi.astype(dtype)
                run_pass_on_host(step, got, n, grid=grid, tile_block=blk)
                assert np.abs(got - want).max() <= tol, (pair, grid, blk)
            # an odd range of tiles: only those tiles change, whatever the mapping
            part = psi.astype(dtype)
            run_pass_on_host(step, part, n, grid=2, tile_range=(3, 10), tile_block=1)
            full = psi.astype(dtype)
            run_pass_on_host(step, full, n, grid=2, tile_block=0)
            changed = np.flatnonzero(np.abs(part - psi.astype(dtype)) > 0)
            assert len(changed) and np.abs(part[changed] - full[changed]).max() <= tol
            same = np.flatnonzero(part == psi.astype(dtype))
            assert len(changed) + len(same) == 1 << n and len(changed) <= 7 * 2048


def test_paired_loads_leave_zero_input_and_scatter_kernels_alone(monkeypatch):
    from tests.jit_host_run import kernel_source
    prog = compile_circuit(W.random_1q_cz(14, 20, 1234))
    assert prog.passes[0].desc.zero_input == 1
    monkeypatch.setenv("QSV_JIT_PAIR", "0")
    plain0, plain1 = kernel_source(prog.passes[0]), kernel_source(prog.passes[1])
    monkeypatch.setenv("QSV_JIT_PAIR", "1")
    assert kernel_source(prog.passes[0]) == plain0          # nothing is read: nothing to pair
    paired = kernel_source(prog.passes[1])
    assert paired != plain1 and "tile1" in paired and "jit_seq(s + 1, F.blk)" in paired and "tile3" not in paired
    monkeypatch.setenv("QSV_JIT_PAIR", "2")
    assert "jit_seq(s + 3, F.blk)" in kernel_source(prog.passes[1])
    assert kernel_source(prog.passes[1], scatter_bits=[12]) == \
        (monkeypatch.setenv("QSV_JIT_PAIR", "0") or kernel_source(prog.passes[1], scatter_bits=[12]))
