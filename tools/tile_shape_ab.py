"""Where does the time of a pass go: arithmetic or access pattern?  Every pass of the n=30 headline program is
timed three times on the same data: (a) as planned, (b) with its tile moved to the LOWEST 11 positions
(one contiguous 32 KB run per tile), (c) with its 8 free tile positions moved to the TOP of the index
(128-byte rows at the largest strides).  The op list, round structure and shared-memory traffic are identical
in all three (the kernels differ only in their address constants), so differences are memory-system effects.
(b) and (c) compute different unitaries than (a) — timing experiment only.   gpurun -- python tools/tile_shape_ab.py"""
import copy
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from quantum_simulations_b200 import _lib as L
from quantum_simulations_b200 import workloads as W
from quantum_simulations_b200.circuit.io import validate_circuit_dict
from quantum_simulations_b200.circuit.sharding import plan_single
from quantum_simulations_b200.kernel.cuda import DeviceState
from quantum_simulations_b200.kernel.cuda_dense import circuit_ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
prog = plan_single(circuit_ops(validate_circuit_dict(W.random_1q_cz(n, 20, 1234))), n, "complex128", True, False)


def remap(step, new_bits):
    """copy of `step` whose tile position i sits on physical bit new_bits[i] (ascending); controls on
    bits outside the tile are dropped (they only gate a few ops)."""
    s = copy.copy(step)
    d = L.QsvPass.from_buffer_copy(step.desc)
    old = list(step.desc.load_bits[: d.n_tile])
    where = {b: i for i, b in enumerate(old)}
    for i in range(d.n_tile):
        d.load_bits[i] = new_bits[i]
        d.store_bits[i] = new_bits[where[step.desc.store_bits[i]]]
    flip = 0
    for i, b in enumerate(old):
        if (step.desc.store_flip >> b) & 1:
            flip |= 1 << new_bits[i]
    d.store_flip = flip
    d.zero_input = 0
    ops = (L.QsvOp * max(step.n_micro_ops, 1))()
    for j in range(step.n_micro_ops):
        ops[j] = step.ops[j]
        ops[j].glob_ctrl = 0
    s.desc, s.ops = d, ops
    return s


rows = []
with DeviceState(n) as st:
    st.init_zero()
    for pi, step in enumerate(prog.passes):
        variants = {"planned": remap(step, list(step.desc.load_bits[:11])),
                    "lowest_11": remap(step, list(range(11))),
                    "low3_top8": remap(step, [0, 1, 2] + list(range(n - 8, n))),
                    "low3_mid8": remap(step, [0, 1, 2] + list(range(11, 19)))}
        row = {"pass": pi, "ops": step.n_micro_ops, "rounds": step.desc.n_rounds, "tile": list(step.desc.load_bits[:11])}
        for name, v in variants.items():
            h = st.upload_steps([v])
            for _ in range(2):
                st.replay(h)
            st.sync()
            st.timer_start()
            for _ in range(5):
                st.replay(h)
            row[name + "_ms"] = round(st.timer_stop() / 5, 3)
            st.release_program(h)
        rows.append(row)
        print(json.dumps(row), flush=True)
