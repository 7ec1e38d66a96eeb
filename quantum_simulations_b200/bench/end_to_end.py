"""End-to-end circuit throughput through the runner API (work dir, WAL, chunk files) with ``kernel="cuda"``.

GPU counterpart of the reference's ``wenbo_engine/bench/end_to_end.py:14-48``: the same two runners
(``runner.single_node.run`` and ``runner.pipeline.run``), the same circuits (GHZ / QFT of its fixtures) and the
same accounting — bytes = 2^n * 8 (complex64 on disk, the reference's unit) * #gates / wall time, MB/s — so the rows
can be read beside the reference's table.  Wall clock (``time.perf_counter``) because this is the user-visible
number: circuit validation, planning, kernel specialisation (cached after the first run), the device work, the
device-to-host copy and the chunk files of the final state are all inside it.

    python -m quantum_simulations_b200.bench.end_to_end [--qubits 20 24] [--json out.jsonl]
"""
from __future__ import annotations

import argparse
import json
import tempfile
import time

from quantum_simulations_b200 import workloads as W
from quantum_simulations_b200.runner.pipeline import run as pl_run
from quantum_simulations_b200.runner.single_node import run as sn_run


def bench_e2e(circ_fn, circ_name: str, chunk_size: int = 0, use_fusion: bool = True, out=None) -> dict:
    cd = circ_fn()
    n = cd["number_of_qubits"]
    N = 1 << n
    if chunk_size == 0:
        chunk_size = N
    total_bytes = N * 8  # complex64 on disk: the reference's unit (bench/end_to_end.py:21)
    results = {}
    for runner_name, runner in [("single_node", sn_run), ("pipeline", pl_run)]:
        with tempfile.TemporaryDirectory() as td:
            t0 = time.perf_counter()
            runner(cd, td, chunk_size=chunk_size, use_fusion=use_fusion)
            dt = time.perf_counter() - t0
        mb_s = total_bytes * len(cd["gates"]) / dt / 1e6
        results[runner_name] = {"time": dt, "MBs": mb_s, "amp_updates_per_s": N * len(cd["gates"]) / dt}
        print(f"  {runner_name:<14} {dt:.4f}s  {mb_s:.1f} MB/s", file=out, flush=True)
    return results


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--qubits", type=int, nargs="*", default=[6, 10, 20, 24])
    ap.add_argument("--no-fusion", action="store_true")
    ap.add_argument("--json", default=None)
    a = ap.parse_args(argv)
    rows = []
    for nq in a.qubits:
        for name, fn in ((f"GHZ-{nq}", lambda nq=nq: W.ghz(nq)), (f"QFT-{nq}", lambda nq=nq: W.qft(nq))):
            print(f"\n{name}  (n={nq}, gates={len(fn()['gates'])})")
            r = bench_e2e(fn, name, use_fusion=not a.no_fusion)
            rows.append({"circuit": name, "n_qubits": nq, "gates": len(fn()["gates"]), **{k: v for k, v in r.items()}})
    if a.json:
        with open(a.json, "a") as f:
            for r in rows:
                f.write(json.dumps(r) + "\n")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
