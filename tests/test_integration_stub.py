"""The ctypes stub printed in INTEGRATION.md is executed as written (only the library path is
filled in): it must bind every symbol it names and, without a GPU, fail loudly instead of falling
back to anything."""
import re
from pathlib import Path

import numpy as np
import pytest

from quantum_simulations_b200 import _lib as L

ROOT = Path(__file__).resolve().parent.parent


def test_documented_stub_binds_and_refuses_without_a_gpu():
    md = (ROOT / "INTEGRATION.md").read_text()
    code = re.search(r"```python\n(\"\"\"kernel=\"cuda\".*?)```", md, re.S).group(1)
    ns: dict = {}
    exec(code.replace("/path/to/libqsv.so", str(L.LIB_PATH)), ns)          # noqa: S102 (our own documentation)
    x = np.zeros(8, dtype=np.complex128)
    x[0] = 1
    X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
    if L.load().qsv_device_count() == 0:
        with pytest.raises(RuntimeError, match="no CUDA device"):
            ns["apply_1q"](x, 0, X)
        assert x[0] == 1                                                     # untouched: no CPU path ran
    else:
        ns["apply_1q"](x, 0, X)
        assert x[1] == 1 and x[0] == 0
        with pytest.raises(NotImplementedError):
            ns["apply_1q"](x, 5, X)
