"""Host side of DeviceState.probabilities: a marginal over few qubits is computed as a wider one (1024 bins, which
the kernel serves without contention) and folded on the host.  The widening and the fold against the oracle's
definition of a marginal, with the device call replaced by the oracle itself."""
import ctypes as C

import numpy as np
import pytest

from oracle import ref_dense as O
from quantum_simulations_b200.kernel import cuda as K


@pytest.mark.parametrize("qubits,n_local,want", [
    ([0], 30, [0, 1, 2, 3, 4, 5, 6, 7, 8, 9]),
    ([29], 30, [29, 0, 1, 2, 3, 4, 5, 6, 7, 8]),
    ([3, 9], 14, [3, 9, 0, 1, 2, 4, 5, 6, 7, 8]),
    ([13, 1, 12], 12, [13, 1, 12, 0, 2, 3, 4, 5, 6, 7]),            # 13, 12: rank bits of a sharded handle; pads are local
    (list(range(12)), 14, list(range(12))),                         # already wide: untouched
    ([], 4, [0, 1, 2, 3]),                                           # small shard: as many as there are
    ([2, 0], 3, [2, 0, 1]),
])
def test_widening(qubits, n_local, want):
    assert K.pad_marginal_qubits(qubits, n_local) == want


class _OracleLib:
    """qsv_probabilities answered by the oracle on a host vector; records what was asked"""
    def __init__(self, psi):
        self.psi, self.asked = psi, []

    def qsv_probabilities(self, h, nq, qs, out):
        qubits = [qs[i] for i in range(nq)]
        self.asked.append(qubits)
        p = O.marginal_probabilities(self.psi, qubits)
        C.memmove(out, p.ctypes.data, p.nbytes)
        return 0


@pytest.mark.parametrize("n", [3, 6, 11, 13])
def test_folded_marginals_are_the_oracles(n):
    rng = np.random.default_rng(n)
    psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    psi /= np.linalg.norm(psi)
    st = K.DeviceState.__new__(K.DeviceState)            # no device: only the host logic of probabilities()
    st.lib, st._h, st.n_local = _OracleLib(psi), None, n
    st._ck = lambda rc: None
    for qs in ([0], [n - 1], [n - 1, 0], [1, 2, 0][: min(3, n)], [], list(range(n))):
        got = st.probabilities(qs)
        assert got.shape == (1 << len(qs),)
        assert np.abs(got - O.marginal_probabilities(psi, qs)).max() <= 1e-14
        asked = st.lib.asked[-1]
        assert asked[: len(qs)] == list(qs) and len(set(asked)) == len(asked) == max(len(qs), min(10, n))
    # a zero-probability outcome stays EXACTLY zero (project() relies on p <= 0 to refuse it)
    basis = np.zeros(1 << n, dtype=np.complex128)
    basis[0] = 1
    st.lib = _OracleLib(basis)
    assert st.probabilities([0]).tolist() == [1.0, 0.0]
