"""Runs bench.py's N>1 control flow on CPU: the GPU objects are replaced by NumPy-emulator fakes
(test infrastructure), the process group is real gloo.  Checks that the JSON line is assembled."""
from __future__ import annotations

import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import quantum_simulations_b200.runner.multi_gpu as MG          # noqa: E402
import quantum_simulations_b200.storage.pinned as PIN            # noqa: E402
from multi_process_worker import EmuShard                                 # noqa: E402


class FakeState:
    def __init__(self, emu):
        self.emu, self._t0, self._timed, self._timing = emu, 0.0, [], False

    def init_zero(self):
        self.emu.psi[:] = 0
        if self.emu.rank == 0:
            self.emu.psi[0] = 1

    def sync(self): pass
    def timing(self, on): self._timing = on
    def timer_start(self): self._t0 = time.perf_counter()
    def timer_stop(self): return (time.perf_counter() - self._t0) * 1e3
    def take_timings(self): out, self._timed = self._timed, []; return out
    def norm2(self): return float(np.vdot(self.emu.psi, self.emu.psi).real)

    def download(self, out=None):
        if out is None:
            return self.emu.psi.copy()
        out[:] = self.emu.psi
        return out


class FakeShard:
    peer_error = "emulated"
    pipelined_swaps = 0

    def __init__(self, emu, state):
        self.emu, self.state = emu, state

    def prepare(self, prog, agree=None): pass
    def release(self, prog=None): pass

    def run_passes(self, steps):
        t0 = time.perf_counter()
        self.emu.run_passes(steps)
        if self.state._timing:
            self.state._timed += [((time.perf_counter() - t0) * 1e3 / len(steps), 10, i) for i in range(len(steps))]

    def swap(self, g, l):
        t0 = time.perf_counter()
        self.emu.swap(g, l)
        if self.state._timing:
            self.state._timed.append(((time.perf_counter() - t0) * 1e3, 20 + len(g), -1))


class FakeSim:
    def __init__(self, n, dtype="complex128"):
        self.rank, self.local_rank, self.world = MG.dist_env()
        self.n, self.g, self.dtype = n, self.world.bit_length() - 1, np.dtype(dtype)
        self.dist = MG.init_plumbing()
        emu = EmuShard(n, self.rank, self.world, self.dist)
        self.shard = FakeShard(emu, FakeState(emu))
        self.peer_swap, self.logical_rank, self._flip_mask = False, self.rank, 0
        self._prepared = {}

    plan = MG.ShardedSimulator.plan
    plan_ops = MG.ShardedSimulator.plan_ops
    simulate_qasm = MG.ShardedSimulator.simulate_qasm
    run = MG.ShardedSimulator.run
    prepare = MG.ShardedSimulator.prepare
    _agree = MG.ShardedSimulator._agree

    def simulate(self, cd, out=None, sink=None, chunk_amps=1 << 24, **kw):
        self.run(self.plan(cd, **kw))
        if sink is not None:
            sink(self.shard.emu.psi, 0)
            return self.shard.emu.psi.nbytes
        got = self.shard.state.download(out)
        if os.environ.get("FAKE_BREAK_PAIRED_LOADS") == "1" and os.environ.get("QSV_JIT_PAIR") != "0":
            got[3] += 1e-3               # stands for a data-movement switch that misbehaves on the sharded path
        if os.environ.get("FAKE_BREAK_PLAN_SEARCH") == "1" and os.environ.get("QSV_PLAN_SEARCH") != "0":
            got[5] += 1e-3               # stands for a searched stage plan that misbehaves
        return got

    def close(self): pass


class FakePinned:
    def __init__(self, nbytes): self.buf = np.zeros(nbytes, dtype=np.uint8)
    def array(self, dtype, count): return self.buf.view(dtype)[:count]
    def free(self): pass


MG.ShardedSimulator = FakeSim
PIN.PinnedBuffer = FakePinned


def _fake_single_gpu_simulate(cd, dtype="complex128", device=0, **kw):
    """stands for kernel.cuda_dense.simulate on rank 0's GPU in the cross-G parity check (test infrastructure)"""
    from oracle import ref_dense as O
    return O.simulate(cd).astype(dtype)


import quantum_simulations_b200.kernel.cuda_dense as CD         # noqa: E402
CD.simulate = _fake_single_gpu_simulate

if __name__ == "__main__":
    import bench
    sys.argv = ["bench.py", "--gpus", sys.argv[1], "--qubits", "12", "--steps", "2", "--warmup", "3",
                "--tile-bits", "6", "--low-bits", "2", "--no-weak"] + (["--no-parity"] if len(sys.argv) < 3 else [])
    if len(sys.argv) >= 3:               # "parity": the parity-first path; the 1-GPU rate needs a device, so it is stubbed
        bench._one_gpu_rate = lambda *a, **k: {"n_qubits": 12, "ms_per_step": 1.0, "value": 1.0, "unit": bench.UNIT, "steps": 1, "warmup": 0}
    bench.main()
