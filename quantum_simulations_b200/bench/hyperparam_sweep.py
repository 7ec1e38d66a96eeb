"""Hyper-parameter sweep over chunk_size, buffer_depth and fusion for the GPU runners.

Mirror of the reference's ``wenbo_engine/bench/hyperparam_sweep.py:33-118`` (same table, same two runners, wall clock
of a whole run without WAL).  On the GPU path the roles of the parameters change, which is what the sweep shows:
``chunk_size`` only sets the size of the checkpoint / result files (every qubit is local to the device), ``buffer_depth``
the number of device snapshots the asynchronous writer may hold, and ``use_fusion`` whether consecutive levels are
planned into shared passes (the one that matters: it divides the number of sweeps over HBM).

    python -m quantum_simulations_b200.bench.hyperparam_sweep [--qubits 20] [--reps 1]
"""
from __future__ import annotations

import argparse
import math
import tempfile
import time

from quantum_simulations_b200 import workloads as W
from quantum_simulations_b200.circuit.fusion import fusion_stats
from quantum_simulations_b200.circuit.io import levelize, validate_circuit_dict
from quantum_simulations_b200.runner.pipeline import run as pl_run
from quantum_simulations_b200.runner.single_node import run as sn_run


def _timed_run(runner, cd, chunk_size, **kwargs):
    with tempfile.TemporaryDirectory() as td:
        t0 = time.perf_counter()
        runner(cd, td, chunk_size=chunk_size, use_wal=False, **kwargs)
        return time.perf_counter() - t0


def sweep(circuit_fn, circuit_name: str, chunk_exponents: list[int] | None = None,
          buffer_depths: list[int] | None = None, reps: int = 1, out=None) -> list[tuple]:
    """Run the sweep and print the results table; returns (runner, chunk exponent, buffer depth, fusion, seconds)."""
    if chunk_exponents is None:
        chunk_exponents = [16, 18, 20]
    if buffer_depths is None:
        buffer_depths = [1, 2, 4, 8]
    cd = validate_circuit_dict(circuit_fn())
    n = cd["number_of_qubits"]
    N = 1 << n
    levels = levelize(cd)
    print(f"\n{'=' * 80}\nHYPERPARAMETER SWEEP: {circuit_name}", file=out)
    print(f"  n_qubits={n}, state_size={N}, gates={len(cd['gates'])}, levels={len(levels)}\n{'=' * 80}\n", file=out)
    print("Fusion analysis (one device: every qubit is local, k = n):", file=out)
    stats = fusion_stats(levels, n)
    print(f"  {stats['io_reduction']}, ops {stats['ops_before']}->{stats['ops_after']}\n", file=out)
    print(f"{'runner':<12} {'chunk':>8} {'buf':>5} {'fusion':>7} {'time(s)':>9} {'speedup':>8}", file=out)
    print("-" * 55, file=out)
    baseline = None
    results = []
    for exp in chunk_exponents:
        cs = min(1 << exp, N)
        for use_fusion in (False, True):
            avg = sum(_timed_run(sn_run, cd, cs, use_fusion=use_fusion) for _ in range(reps)) / reps
            baseline = avg if baseline is None else baseline
            tag = "yes" if use_fusion else "no"
            print(f"{'single_node':<12} {f'2^{int(math.log2(cs))}':>8} {'--':>5} {tag:>7} {avg:>9.4f} {baseline / avg:>7.2f}x", file=out, flush=True)
            results.append(("single_node", exp, 0, use_fusion, avg))
            for bd in buffer_depths:
                avg = sum(_timed_run(pl_run, cd, cs, buffer_depth=bd, use_fusion=use_fusion) for _ in range(reps)) / reps
                print(f"{'pipeline':<12} {f'2^{int(math.log2(cs))}':>8} {bd:>5} {tag:>7} {avg:>9.4f} {baseline / avg:>7.2f}x", file=out, flush=True)
                results.append(("pipeline", exp, bd, use_fusion, avg))
    best = min(results, key=lambda r: r[4])
    print(f"\nbest: {best[0]} chunk=2^{best[1]} buffer_depth={best[2] or '--'} fusion={'yes' if best[3] else 'no'}: {best[4]:.4f}s", file=out)
    return results


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--qubits", type=int, default=20)
    ap.add_argument("--reps", type=int, default=1)
    a = ap.parse_args(argv)
    sweep(lambda: W.qft(a.qubits), f"QFT-{a.qubits}", reps=a.reps)
    sweep(lambda: W.ghz(a.qubits), f"GHZ-{a.qubits}", reps=a.reps)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
