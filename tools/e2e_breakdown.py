"""Where the end-to-end time of kernel.cuda_dense.simulate goes (run on the GPU box)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from quantum_simulations_b200 import workloads as W
from quantum_simulations_b200.circuit.io import validate_circuit_dict
from quantum_simulations_b200.circuit.passes import PassCompiler
from quantum_simulations_b200.kernel.cuda import DeviceState
from quantum_simulations_b200.kernel.cuda_dense import circuit_ops
from quantum_simulations_b200.storage.pinned import PinnedBuffer

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
host = PinnedBuffer((1 << n) * 16)
out = host.array("complex128", 1 << n)
for rep in range(3):
    t = [time.perf_counter()]
    cd = validate_circuit_dict(W.random_1q_cz(n, 20, 1234)); ops = circuit_ops(cd); t.append(time.perf_counter())
    st = DeviceState(n); t.append(time.perf_counter())
    st.init_zero(); st.sync(); t.append(time.perf_counter())
    prog = PassCompiler(n).compile(ops); t.append(time.perf_counter())
    h = st.upload_steps(prog.steps); t.append(time.perf_counter())
    st.replay(h); st.sync(); t.append(time.perf_counter())
    st.download(out); t.append(time.perf_counter())
    st.close(); t.append(time.perf_counter())
    names = ["validate+matrices", "create(cudaMalloc)", "init_zero", "pass compiler", "program_create(jit lookup)", "replay", "download D2H", "close(cudaFree)"]
    print(rep, {k: round((b - a) * 1e3, 1) for k, a, b in zip(names, t, t[1:])}, "total", round((t[-1] - t[0]) * 1e3, 1))
host.free()
