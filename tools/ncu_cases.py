"""A few fixed pass-kernel launches for an ncu capture (run under gpurun + ncu)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from quantum_simulations_b200 import _lib as L
from quantum_simulations_b200 import workloads as W
from quantum_simulations_b200.kernel.cuda import DeviceState
from quantum_simulations_b200.kernel.cuda_dense import compile_circuit
from tools.sweep_pass import make_pass

n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
with DeviceState(n) as st:
    st.init_zero()
    prog = compile_circuit(W.random_1q_cz(n, 20, 1234), low_bits=3)
    real_pass = max(prog.passes, key=lambda s: s.n_micro_ops)
    cases = [make_pass(n, 11, 5, 1, 0), make_pass(n, 11, 5, 4, 0),
             make_pass(n, 11, 5, 1, 16, kind=L.OP_HAD), make_pass(n, 11, 5, 1, 16, kind="SIGN_FOLD"),
             real_pass]
    for c in cases:          # warm-up launches (skipped by ncu -s)
        st.apply_pass(c)
    st.sync()
    for c in cases:
        st.apply_pass(c)
    st.sync()
print("ok", real_pass.n_micro_ops, real_pass.desc.n_rounds)
