"""Runs bench.py's N = 1 control flow on CPU: DeviceState / simulate / PinnedBuffer are replaced by
NumPy-emulator fakes (test infrastructure).  Checks that the JSON line is assembled without a GPU."""
from __future__ import annotations

import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import quantum_simulations_b200.kernel.cuda as KC                 # noqa: E402
import quantum_simulations_b200.kernel.cuda_dense as KD           # noqa: E402
import quantum_simulations_b200.storage.pinned as PIN             # noqa: E402
from quantum_simulations_b200.circuit.io import validate_circuit_dict   # noqa: E402
from pass_emulator import run_program                             # noqa: E402


class FakeState:
    def __init__(self, n, dtype="complex128", device=0, rank=0, world=1):
        self.psi = np.zeros(1 << n, dtype=np.complex128)
        self._timed, self._timing, self._t0 = [], False, 0.0

    def __enter__(self): return self
    def __exit__(self, *exc): pass
    def upload_program(self, prog): return prog
    def init_zero(self): self.psi[:] = 0; self.psi[0] = 1
    def sync(self): pass
    def timing(self, on): self._timing = on
    def timer_start(self): self._t0 = time.perf_counter()
    def timer_stop(self): return (time.perf_counter() - self._t0) * 1e3
    def take_timings(self): out, self._timed = self._timed, []; return out
    def norm2(self): return float(np.vdot(self.psi, self.psi).real)

    def replay(self, prog):
        if prog.fused_init:
            self.init_zero()                       # what the zero-input pass stands for
        t0 = time.perf_counter()
        run_program(prog, self.psi)
        if self._timing:
            k = len(prog.passes)
            self._timed += [((time.perf_counter() - t0) * 1e3 / k, 10, i) for i in range(k)]

    def download(self, out=None, offset=0, count=None):
        if out is None:
            return self.psi[offset: None if count is None else offset + count].copy()
        out[:] = self.psi
        return out


def fake_simulate(cd, dtype="complex128", device=0, out=None, phases=None, **kw):
    cd = validate_circuit_dict(cd)
    kw.pop("skip_zero_support", None)
    prog = KD.compile_circuit(cd, dtype=dtype, **kw)
    st = FakeState(cd["number_of_qubits"])
    st.replay(prog) if prog.fused_init else (st.init_zero(), st.replay(prog))
    if phases is not None:
        phases["run"] = phases.get("run", 0.0) + 1.0
    return st.download(out)


class FakePinned:
    def __init__(self, nbytes): self.buf = np.zeros(nbytes, dtype=np.uint8)
    def array(self, dtype, count): return self.buf.view(dtype)[:count]
    def free(self): pass


KC.DeviceState = FakeState
KD.simulate = fake_simulate
PIN.PinnedBuffer = FakePinned

if __name__ == "__main__":
    import bench
    bench.jit_stats = lambda: {"kernels_compiled": 0}
    sys.argv = ["bench.py", "--qubits", "12", "--steps", "2", "--warmup", "3", "--tile-bits", "6", "--low-bits", "2",
                "--cpu-qubits", "10"]
    bench.main()
