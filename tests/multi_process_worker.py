"""Worker of tests/test_multi_process_host.py: one PROCESS per shard, the host plumbing (runner/plumbing.py,
plain TCP) for the exchange.

Runs the product-side multi-GPU logic (sharding.plan + runner.multi_gpu.execute) with the NumPy
pass emulator standing in for libqsv (there is no GPU here); the SwapStep is a real
all-to-all between processes with the same block arithmetic as csrc/exchange.cuh."""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


class EmuShard:
    def __init__(self, n, rank, world, dist):
        import math
        self.n, self.rank, self.world, self.dist = n, rank, world, dist
        self.n_local = n - int(math.log2(world))
        self.psi = np.zeros(1 << self.n_local, dtype=np.complex128)
        if rank == 0:
            self.psi[0] = 1.0
        self.swaps = 0

    def run_passes(self, steps):
        from tests.pass_emulator import run_pass
        for s in steps:
            run_pass(self.psi, s.desc, s.ops, self.n_local, self.rank, s.tables)

    def swap(self, global_bits, local_bits):
        s = len(global_bits)
        assert list(local_bits) == [self.n_local - s + i for i in range(s)]
        me = 0
        for i, gb in enumerate(global_bits):
            me |= ((self.rank >> (gb - self.n_local)) & 1) << i
        blocks = self.psi.reshape(1 << s, -1)
        out = {}
        for d in range(1 << s):
            if d == me:
                continue
            peer = self.rank
            for i, gb in enumerate(global_bits):
                rb = gb - self.n_local
                peer = (peer & ~(1 << rb)) | (((d >> i) & 1) << rb)
            out[peer] = (d, blocks[d].copy())
        box = self.dist.all_gather_object(out)          # every rank: {destination rank: (block index there... , data)}
        for src, sent in enumerate(box):
            if self.rank in sent:
                _, data = sent[self.rank]
                # the block I receive from `src` replaces my block with the index of src's swapped bits
                d_src = 0
                for i, gb in enumerate(global_bits):
                    d_src |= ((src >> (gb - self.n_local)) & 1) << i
                blocks[d_src] = data
        self.swaps += 1


def main():
    from quantum_simulations_b200 import workloads as W
    from quantum_simulations_b200.circuit import sharding
    from quantum_simulations_b200.circuit.io import validate_circuit_dict
    from quantum_simulations_b200.kernel.cuda_dense import circuit_ops
    from quantum_simulations_b200.runner.multi_gpu import execute, dist_env, init_plumbing

    rank, _, world = dist_env()
    dist = init_plumbing()
    out_dir = Path(sys.argv[1])
    n = int(sys.argv[2])
    for name, cd in (("random_1q_cz", W.random_1q_cz(n, 12, 99)), ("qft", W.qft(n)), ("random_mixed", W.random_mixed(n, 100, 4))):
        cd = validate_circuit_dict(cd)
        g = world.bit_length() - 1
        prog = sharding.plan(circuit_ops(cd), n, n - g, tile_bits=6, low_bits=2)
        shard = EmuShard(n, rank, world, dist)
        execute(prog, shard)
        np.save(out_dir / f"{name}_rank{rank}.npy", shard.psi)
        if rank == 0:
            (out_dir / f"{name}_swaps.txt").write_text(str(shard.swaps))
    # the OpenQASM front end on a sharded state (what ShardedSimulator.simulate_qasm does): compiled ZZ
    # blocks behind a Hadamard layer fuse into diagonals, the rest plans into passes and swaps
    from quantum_simulations_b200.circuit.fusion import fuse_2q_blocks
    from quantum_simulations_b200.circuit.qasm import qasm_to_ops
    nq, ops = qasm_to_ops(qasm_text(n))
    g = world.bit_length() - 1
    prog = sharding.plan(fuse_2q_blocks(ops, tol=1e-14), nq, nq - g, tile_bits=6, low_bits=2, rank_flips=True)
    shard = EmuShard(nq, rank, world, dist)
    execute(prog, shard)
    np.save(out_dir / f"qasm_rank{rank ^ prog.rank_flip_mask}.npy", shard.psi)
    dist.barrier()
    dist.close()


def qasm_text(n: int) -> str:
    src = ["OPENQASM 2.0;", 'include "qelib1.inc";', f"qreg q[{n}];", f"h q;"]
    for layer in range(3):
        for a in range(layer % 2, n - 1, 2):
            b = a + 1
            src += [f"rz(0.2) q[{a}];", f"cx q[{a}],q[{b}];", f"rz({0.3 + 0.1 * a}) q[{b}];", f"cx q[{a}],q[{b}];"]
        src += [f"rx({0.4 + 0.05 * q}) q[{q}];" for q in range(n)]
    src += [f"ryy(0.7) q[{n - 1}],q[0];", f"ccx q[0],q[{n // 2}],q[{n - 1}];"]
    return "\n".join(src)


if __name__ == "__main__":
    main()
