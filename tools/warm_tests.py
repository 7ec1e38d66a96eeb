"""Pre-build the specialised kernels of the GPU parity tests named below (NVRTC, no GPU needed), so that a short
GPU call spends its seconds on running them."""
import ctypes as C
import os
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from quantum_simulations_b200 import _lib, workloads as W
from quantum_simulations_b200.circuit.io import validate_circuit_dict
from quantum_simulations_b200.circuit.sharding import plan_single
from quantum_simulations_b200.kernel.cuda_dense import circuit_ops

lib = _lib.load()
os.environ["QSV_JIT_WARM"] = "1"
jobs = []
for n, zs in ((18, False), (20, True)):          # test_specialised_and_interpreted..., test_zero_support_skipping...
    for cd in (W.random_1q_cz(n, 20, 1234), W.random_mixed(n, 400, 5), W.qft(n), W.ghz(n)):
        cd = validate_circuit_dict(cd)
        for dtype in ("complex128", "complex64"):
            jobs += [(s, _lib.QSV_C128 if dtype == "complex128" else _lib.QSV_C64)
                     for s in plan_single(circuit_ops(cd), n, dtype, True, zs).passes]
for cd, dtype in ((W.random_1q_cz(30, 20, 1234), "complex64"), (W.random_1q_cz(26, 20, 1234), "complex128"), (W.qft(12), "complex64")):
    cd = validate_circuit_dict(cd)
    jobs += [(s, _lib.QSV_C128 if dtype == "complex128" else _lib.QSV_C64)
             for s in plan_single(circuit_ops(cd), cd["number_of_qubits"], dtype, True, False).passes]


def one(job):
    step, dt = job
    n_, log = C.c_size_t(), C.create_string_buffer(4096)
    return lib.qsv_jit_build_pass(C.byref(step.desc), step.ops, dt, C.byref(n_), log, len(log))


for env in sys.argv[1:] or [""]:
    if env:
        os.environ["QSV_JIT_PAIR"] = env
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        rcs = list(ex.map(one, jobs))
    print(f"QSV_JIT_PAIR={env or 'default'}: {sum(r == 0 for r in rcs)}/{len(jobs)} kernels built")
