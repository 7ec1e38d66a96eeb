"""Worker of tests/test_multi_gpu.py (run under torchrun on a multi-GPU box): real libqsv
shards + NVLink swaps, gathered over the host plumbing on rank 0 and compared with the C oracle."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    from oracle import c_oracle as CO
    from quantum_simulations_b200 import workloads as W
    from quantum_simulations_b200.circuit.io import validate_circuit_dict
    from quantum_simulations_b200.runner.multi_gpu import ShardedSimulator

    n = int(sys.argv[1])
    sim = ShardedSimulator(n)
    dist, rank, world = sim.dist, sim.rank, sim.world
    worst = 0.0
    g = world.bit_length() - 1
    # work on the low qubits, then mix every rank-bit qubit: the last pass before the swap leaves the
    # top local bits alone, so the exchange must run OVERLAPPED with it (qsv_pass_swap_overlapped)
    low_then_top = {"number_of_qubits": n, "gates":
                    [{"qubits": [q], "gate": "RY", "params": {"theta": 0.3 + 0.1 * q}} for q in range(10)]
                    + [{"qubits": [q, q + 1], "gate": "CZ"} for q in range(9)]
                    + [{"qubits": [q], "gate": "H"} for q in range(10)]
                    + [{"qubits": [q], "gate": "H"} for q in range(n - g, n)]
                    + [{"qubits": [n - 1, 0], "gate": "CNOT"}]}
    overlapped_seen = 0
    for name, cd in (("low_then_top", low_then_top), ("random_1q_cz", W.random_1q_cz(n, 20, 1234)), ("qft", W.qft(n)),
                     ("ghz", W.ghz(n)), ("random_mixed", W.random_mixed(n, 300, 8))):
        cd = validate_circuit_dict(cd)
        if name == "low_then_top":
            # identity placement (as after an upload / resume): the planner cannot dodge the swap
            from quantum_simulations_b200.circuit.passes import PassCompiler
            from quantum_simulations_b200.kernel.cuda_dense import circuit_ops
            prog = PassCompiler(n, n - g, swap_anywhere=sim.peer_swap, rank_flips=True).compile(circuit_ops(cd))
            sim.run(prog)
            shard = sim.shard.state.download()
        else:
            shard = sim.simulate(cd)
        samples = sim.sample(seed=7, shots=257)
        parts = dist.all_gather_object(shard)
        if rank == 0:
            mask = sim.logical_rank ^ sim.rank           # physical shard r holds logical shard r ^ mask
            got = np.concatenate([parts[l ^ mask] for l in range(world)])
            want = CO.simulate_c(cd)
            err = float(np.abs(got - want).max())
            print(f"{name}: n={n} world={world} max|d|={err:.3e} swaps={sim.shard.swaps} pipelined={sim.shard.pipelined_swaps} "
                  f"scatter_passes={sim.shard.fused_swaps}", flush=True)
            worst = max(worst, err)
            if name == "low_then_top":
                overlapped_seen = sim.shard.pipelined_swaps + sim.shard.fused_swaps
            from oracle import ref_dense as O
            want_s = O.sample_indices(got, 7, 257)
            if not np.array_equal(samples, want_s):
                # the state differs from the oracle's in the last bits, so compare with the samples of
                # the state this run produced: the definition is what must be reproduced bit-exactly
                raise SystemExit(f"{name}: sharded samples differ from oracle.sample_indices on the same state")
    sim.close()
    dist.barrier()
    # the runner surface: chunk files + manifest + WAL written by all ranks, read back by collect_state
    import tempfile, os
    from quantum_simulations_b200.runner import multi_gpu as MG
    from quantum_simulations_b200.runner.single_node import collect_state
    box = [dist.broadcast_object(tempfile.mkdtemp(prefix="qsv_mg_") if rank == 0 else None, src=0)]
    cd = validate_circuit_dict(W.random_1q_cz(n, 10, 21))
    buf = MG.run(cd, box[0], chunk_size=1 << (n - 5), dtype="complex128")
    if rank == 0:
        got = collect_state(buf)
        err = float(np.abs(got - CO.simulate_c(cd)).max())
        print(f"runner.multi_gpu.run: {len(os.listdir(buf / 'chunks'))} chunk files, max|d|={err:.3e}", flush=True)
        worst = max(worst, err)
    dist.barrier()
    # resumable checkpoints: segments of 4 levels; the first call stops after ONE checkpoint (as a crash would),
    # the second resumes from the WAL's done_steps and finishes; a third call on the finished directory returns at once
    box = [dist.broadcast_object(tempfile.mkdtemp(prefix="qsv_mg_ck_") if rank == 0 else None, src=0)]
    cd = validate_circuit_dict(W.random_1q_cz(n, 12, 33))
    import json as _json
    MG.run(cd, box[0], chunk_size=1 << (n - 5), checkpoint_every=4, stop_after_checkpoints=1)
    done_first = _json.loads((Path(box[0]) / "wal.json").read_text())["done_steps"] if rank == 0 else None
    buf = MG.run(cd, box[0], chunk_size=1 << (n - 5), checkpoint_every=4)
    buf_again = MG.run(cd, box[0], chunk_size=1 << (n - 5), checkpoint_every=4)
    if rank == 0:
        wal = _json.loads((Path(box[0]) / "wal.json").read_text())
        got = collect_state(buf)
        err = float(np.abs(got - CO.simulate_c(cd)).max())
        print(f"runner.multi_gpu.run resumed: done_steps {done_first} -> {wal['done_steps']} of 12 levels, max|d|={err:.3e}", flush=True)
        if done_first != 4 or wal["done_steps"] != 12 or str(buf_again) != str(buf):
            raise SystemExit(f"resume bookkeeping wrong: first={done_first} wal={wal} again={buf_again} vs {buf}")
        worst = max(worst, err)
    dist.barrier()
    # use_staging=True: the reference's atlas_stages supplies the stages; chunks in the physical layout +
    # qubit_mapping.json, un-permuted by collect_state exactly like a reference work directory
    box = [dist.broadcast_object(tempfile.mkdtemp(prefix="qsv_mg_st_") if rank == 0 else None, src=0)]
    cd = validate_circuit_dict(W.random_mixed(n, 200, 17))
    for method in ("heuristic", "ilp"):
        buf = MG.run(cd, str(Path(box[0]) / method), chunk_size=1 << (n - 5), use_staging=True, staging_method=method)
        if rank == 0:
            got = collect_state(buf, apply_permutation=True, work_dir=str(Path(box[0]) / method))
            err = float(np.abs(got - CO.simulate_c(cd)).max())
            l2p = _json.loads((Path(box[0]) / method / "qubit_mapping.json").read_text())
            print(f"runner.multi_gpu.run use_staging={method}: mapping identity={l2p == list(range(n))} max|d|={err:.3e}", flush=True)
            worst = max(worst, err)
        dist.barrier()
    dist.close()
    if rank == 0 and worst > 1e-12:
        raise SystemExit(f"multi-GPU parity failed: {worst}")
    if rank == 0 and sim.peer_swap and overlapped_seen < 1:
        raise SystemExit("the overlapped pass+swap path (or, with QSV_FUSED_EXCHANGE=1, the scatter pass) was never taken")
    if rank == 0 and os.environ.get("QSV_FUSED_EXCHANGE") == "1" and not sim.fused_exchange:
        raise SystemExit(f"QSV_FUSED_EXCHANGE=1 but the second buffers could not be wired: {sim.shard.fused_error}")


if __name__ == "__main__":
    main()
