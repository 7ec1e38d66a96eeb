"""Amplitude-table form of a state, as the reference's SQL / Spark generations store it
(v1_implementation, v2_spark/src/gate_applicator.py:155-372, v3_hisvsim_spark: rows
``(idx BIGINT, real DOUBLE, imag DOUBLE)``, entries with |re| and |im| <= 1e-15 dropped;
SURVEY.md section 8a row a14, section 8f-4).  Converters so that states written by those
implementations can be uploaded into the engine (or compared with its output)."""
from __future__ import annotations

from pathlib import Path

import numpy as np

DROP_TOL = 1e-15        # the reference's threshold for omitting a row


def rows_to_dense(idx, real, imag, n_qubits: int, dtype=np.complex128) -> np.ndarray:
    """Dense 2^n vector from parallel arrays of a sparse amplitude table (duplicates add up, like the
    reference's ``GROUP BY idx`` / ``sum``)."""
    idx = np.asarray(idx, dtype=np.int64)
    if idx.size and (idx.min() < 0 or idx.max() >= (1 << n_qubits)):
        raise ValueError("row index outside [0, 2^n)")
    out = np.zeros(1 << n_qubits, dtype=np.complex128)
    np.add.at(out, idx, np.asarray(real, dtype=np.float64) + 1j * np.asarray(imag, dtype=np.float64))
    return out.astype(dtype, copy=False)


def dense_to_rows(psi: np.ndarray, tol: float = DROP_TOL):
    """(idx, real, imag) of the entries the reference would keep (|re| > tol or |im| > tol)."""
    psi = np.asarray(psi)
    keep = np.nonzero((np.abs(psi.real) > tol) | (np.abs(psi.imag) > tol))[0]
    return keep.astype(np.int64), psi.real[keep].astype(np.float64), psi.imag[keep].astype(np.float64)


def read_table(path: str | Path, n_qubits: int, dtype=np.complex128) -> np.ndarray:
    """Dense state from a Parquet file / directory (Spark output) or a CSV with the columns idx, real, imag."""
    path = Path(path)
    if path.suffix.lower() == ".csv":
        data = np.genfromtxt(path, delimiter=",", names=True)
        data = np.atleast_1d(data)
        return rows_to_dense(data["idx"], data["real"], data["imag"], n_qubits, dtype)
    import pyarrow.parquet as pq
    t = pq.read_table(str(path), columns=["idx", "real", "imag"])
    return rows_to_dense(t["idx"].to_numpy(), t["real"].to_numpy(), t["imag"].to_numpy(), n_qubits, dtype)


def write_table(path: str | Path, psi: np.ndarray, tol: float = DROP_TOL) -> int:
    """Write the kept rows as Parquet (or CSV by suffix); returns the number of rows."""
    idx, re, im = dense_to_rows(psi, tol)
    path = Path(path)
    if path.suffix.lower() == ".csv":
        np.savetxt(path, np.column_stack([idx, re, im]), delimiter=",", header="idx,real,imag", comments="",
                   fmt=["%d", "%.17g", "%.17g"])
    else:
        import pyarrow as pa
        import pyarrow.parquet as pq
        pq.write_table(pa.table({"idx": idx, "real": re, "imag": im}), str(path))
    return len(idx)
