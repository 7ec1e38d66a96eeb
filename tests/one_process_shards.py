"""Worker of tests/test_pipelined_swap.py: all shards of a sharded run as handles of THIS process on ONE
device (qsv_comm_init_local / qsv_comm_set_peers_local), so that the sharded execution paths — plain
swaps on the exchange kernels and PIPELINED stage transitions (qsv_swap_pipelined) — are checked against
the oracle on a single-GPU box.  Run as a subprocess: CUDA_DEVICE_MAX_CONNECTIONS must be set before CUDA
starts (a spinning exchange kernel must not share a hardware queue with the stream it waits for)."""
from __future__ import annotations

import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def run_case(n, world, dtype, cd, pipeline, xchg_sms=3, planner_kw=None, parts=False):
    from quantum_simulations_b200.circuit import sharding
    from quantum_simulations_b200.circuit.io import validate_circuit_dict
    from quantum_simulations_b200.kernel.cuda_dense import circuit_ops
    from quantum_simulations_b200.runner.multi_gpu import CudaShard, execute

    g = world.bit_length() - 1
    cd = validate_circuit_dict(cd)
    if parts:     # HiSVSIM execution model: the circuit cut into parts, every part one stage (sharding.plan_parts)
        ops = circuit_ops(cd)
        prog = sharding.plan_parts(sharding.split_into_parts(ops, n - g - 2), n, n - g, dtype, swap_anywhere=True)
    else:
        prog = sharding.plan(circuit_ops(cd), n, n - g, dtype, swap_anywhere=True, rank_flips=True, **(planner_kw or {}))
    shards = [CudaShard(n, r, world, dtype, device=0, local=True) for r in range(world)]
    try:
        CudaShard.wire_local(shards)
        for s in shards:
            s.pipeline = pipeline
            s.xchg_sms = xchg_sms
            s.prepare(prog, agree=lambda f: f)          # same device, same kernels: the shards agree by construction
        n_tr = len(shards[0]._transitions)
        for s in shards:                                 # everything is enqueued asynchronously, rank after rank
            if not prog.fused_init:
                s.state.init_zero()
            execute(prog, s)
        for s in shards:
            s.state.sync()
        got = np.concatenate([shards[l ^ prog.rank_flip_mask].state.download() for l in range(world)])
        info = {"swaps": prog.stats["swaps"], "transitions": n_tr, "pipelined": shards[0].pipelined_swaps,
                "plans": [[t.a_count, t.b_count, t.chunk_bits] for _, t in shards[0]._transitions.values()]}
    finally:
        for s in shards:
            s.close()
    return got, info


def main():
    from oracle import ref_dense as O
    from quantum_simulations_b200 import workloads as W
    from quantum_simulations_b200.circuit.io import validate_circuit_dict

    import signal
    signal.alarm(150)            # a deadlocked exchange must not hold the GPU box: die instead
    cases = json.loads(sys.argv[1])
    worst = 0.0
    for case in cases:
        n, world, dtype, name = case["n"], case["world"], case["dtype"], case["circuit"]
        cd = {"random": lambda: W.random_1q_cz(n, 20, 1234), "qft": lambda: W.qft(n), "ghz": lambda: W.ghz(n),
              "mixed": lambda: W.random_mixed(n, 300, 8)}[name]()
        want = O.simulate(validate_circuit_dict(cd))
        for pipeline in (True, False):
            got, info = run_case(n, world, dtype, cd, pipeline, case.get("xchg_sms", 3), parts=case.get("parts", False))
            err = float(np.abs(got - want.astype(got.dtype)).max())
            print(json.dumps({"case": case, "pipeline": pipeline, "max_abs_err": err, **info}), flush=True)
            tol = 1e-12 if dtype == "complex128" else 2e-5
            if err > tol:
                raise SystemExit(f"parity failed: {case} pipeline={pipeline} err={err}")
            if pipeline and case.get("expect_pipelined", True) and info["pipelined"] < 1:
                raise SystemExit(f"no pipelined transition ran for {case}: {info}")
            worst = max(worst, err)
    print(json.dumps({"ok": True, "worst": worst}))


if __name__ == "__main__":
    main()
