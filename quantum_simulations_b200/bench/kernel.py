"""Per-gate kernel and observable throughput on the device, against the HBM roofline.

GPU counterpart of the reference's ``wenbo_engine/bench/kernel.py:11-54`` (scalar vs batched GB/s of one gate on a
chunk): every per-gate entry point of the C ABI (``qsv_apply_1q / 2q / ctrl_1q / diag / kq``, the a1/a2 operator
face, csrc/gate_kernels.cuh) is timed on a state resident in HBM for target qubits at low, middle and high index
bits, and the read-only observable kernels (``qsv_norm2 / probabilities / expect_z / sample``, csrc/sample.cuh)
beside them.  Timing: CUDA events on the handle's stream (``qsv_timer_start/stop``), after a warm-up launch.

Accounting: a gate kernel reads and writes every amplitude it touches once, so its ALGORITHMIC traffic is
2 * sizeof(amp) * (amplitudes touched) per launch (a controlled gate touches half of them, a doubly controlled
quarter); an observable reads the state once.  ``gbs`` = that traffic / time; ``frac`` = gbs / the measured HBM
copy peak (MEASURED_PEAKS.json ``hbm_gbs``; 6555.2 GB/s if the file is absent).  ``ref_gbs`` is the reference's
own accounting (chunk bytes * reps / time: one touch per amplitude, bench/kernel.py:21) for side-by-side reading.

    python -m quantum_simulations_b200.bench.kernel [n_qubits] [--dtype complex64] [--reps 5] [--json out.jsonl]
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import numpy as np

from quantum_simulations_b200.kernel import gates as gmod
from quantum_simulations_b200.kernel.cuda import DeviceState

_FALLBACK_PEAK_GBS = 6555.2


def hbm_peak_gbs() -> tuple[float, str]:
    p = Path(__file__).resolve().parents[2] / "MEASURED_PEAKS.json"
    try:
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return _FALLBACK_PEAK_GBS, "fallback 6555.2 GB/s (MEASURED_PEAKS.json absent)"


def _time(st: DeviceState, fn, reps: int) -> float:
    """mean milliseconds of one call of fn, CUDA events around `reps` back-to-back calls after one warm-up"""
    fn()
    st.sync()
    st.timer_start()
    for _ in range(reps):
        fn()
    return st.timer_stop() / reps


def _row(st: DeviceState, name: str, what: str, touched: float, writes: bool, ms: float, peak: float) -> dict:
    amp = st.dtype.itemsize
    traffic = (2 if writes else 1) * amp * touched * st.n_amps
    gbs = traffic / ms / 1e6
    return {"kernel": name, "what": what, "n_qubits": st.n_qubits, "dtype": st.dtype.name, "ms": round(ms, 4),
            "algorithmic_bytes": int(traffic), "gbs": round(gbs, 1), "frac": round(gbs / peak, 4),
            "ref_gbs": round(amp * st.n_amps / ms / 1e6, 1)}


def _bench_1q(st: DeviceState, qubit: int, gate_name: str = "H", reps: int = 5) -> float:
    U = gmod.gate_matrix(gate_name, {})
    return _time(st, lambda: st.apply_1q(qubit, U), reps)


def _bench_2q(st: DeviceState, qa: int, qb: int, gate_name: str = "CNOT", reps: int = 5) -> float:
    U = gmod.gate_matrix(gate_name, {})
    return _time(st, lambda: st.apply_2q(qa, qb, U), reps)


def _spread(n: int) -> list[int]:
    """target qubits worth timing: the low bits (inside one 128-byte line), one in the middle, the top ones"""
    return sorted({q for q in (0, 1, 2, 3, 4, 7, n // 2, n - 2, n - 1) if 0 <= q < n})


def bench_kernel(n_qubits: int = 30, dtype: str = "complex128", reps: int = 5, device: int = 0,
                 observables: bool = True, out=sys.stdout, quick: bool = False) -> list[dict]:
    """quick=True: one representative case per kernel (what `bench.py` puts into its line as `per_gate_kernels`)."""
    peak, peak_src = hbm_peak_gbs()
    rows: list[dict] = []
    n = n_qubits
    rng = np.random.default_rng(7)
    with DeviceState(n, dtype, device) as st:
        st.init_zero()
        # a state with every amplitude non-zero, so that no kernel sees a trivially sparse input
        H = gmod.gate_matrix("H", {})
        for q in range(n):
            st.apply_1q(q, H)

        def add(name, what, touched, writes, ms):
            rows.append(_row(st, name, what, touched, writes, ms, peak))
            r = rows[-1]
            print(f"{r['kernel']:<16} {r['what']:<34} {r['ms']:>8.3f} ms {r['gbs']:>8.1f} GB/s  {r['frac']:>6.3f}", file=out, flush=True)

        print(f"n = {n} ({st.dtype.name}, {st.n_amps * st.dtype.itemsize / 2**30:.2f} GiB), HBM peak {peak:.1f} GB/s: {peak_src}", file=out)
        print(f"{'kernel':<16} {'what':<34} {'time':>11} {'traffic':>13}  {'frac':>6}", file=out)
        for q in ((0, n // 2, n - 1) if quick else _spread(n)):
            add("k_apply_1q", f"H on qubit {q}", 1.0, True, _bench_1q(st, q, "H", reps))
        for g in (() if quick else ("X", "T")):   # the reference's other two 1-qubit cases (bench/kernel.py:46)
            add("k_apply_1q", f"{g} on qubit 0 (dense 2x2 path)", 1.0, True, _bench_1q(st, 0, g, reps))
        pairs = [(0, 1), (n - 2, n - 1)] if quick else [(0, 1), (1, 0), (0, n - 1), (n // 2, n // 2 + 1), (n - 2, n - 1)]
        for qa, qb in pairs:
            add("k_apply_2q", f"CNOT on ({qa},{qb}) dense 4x4", 1.0, True, _bench_2q(st, qa, qb, "CNOT", reps))
        ry = gmod.gate_matrix("RY", {"theta": 0.37})
        for c, t in (((0, n - 1), (n - 1, 0)) if quick else ((1, 0), (0, n - 1), (n - 1, 0), (n // 2, n // 2 + 1))):
            add("k_apply_ctrl_1q", f"controlled RY, ctrl {c} tgt {t}", 0.5, True,
                _time(st, lambda c=c, t=t: st.apply_ctrl_1q(c, t, ry), reps))
        for qs in (((0, n - 1),) if quick else ((0,), (n - 1,), (0, 1), (0, n - 1), (0, 1, 2, n // 2, n - 2, n - 1))):
            ph = np.exp(1j * rng.uniform(0, 2 * np.pi, 1 << len(qs)))
            add("k_apply_diag", f"diagonal on {len(qs)} qubit(s) {list(qs)}"[:34], 1.0, True,
                _time(st, lambda qs=qs, ph=ph: st.apply_diag(list(qs), ph), reps))
        for qs in (((n - 3, n - 2, n - 1), (1, 5, n // 2, n - 4, n - 1)) if quick else
                   ((0, 1, 2), (n - 3, n - 2, n - 1), (0, 1, 2, 3, 4), (1, 5, n // 2, n - 4, n - 1))):
            k = len(qs)
            m = rng.normal(size=(1 << k, 1 << k)) + 1j * rng.normal(size=(1 << k, 1 << k))
            U, _ = np.linalg.qr(m)
            add("k_apply_kq", f"dense {k}-qubit unitary {list(qs)}"[:34], 1.0, True,
                _time(st, lambda qs=qs, U=U: st.apply_kq(list(qs), U), reps))
        if observables:
            add("k_norm2_partial", "sum |amp|^2", 1.0, False, _time(st, st.norm2, reps))
            for qs in (((0,),) if quick else ((0,), (n - 1,), tuple(range(min(10, n))))):
                add("probabilities", f"marginal over {len(qs)} qubit(s) from {qs[0]}", 1.0, False,
                    _time(st, lambda qs=qs: st.probabilities(list(qs)), reps))
            add("expect_z", f"<Z_0 Z_{n - 1}>", 1.0, False, _time(st, lambda: st.expect_z([0, n - 1]), reps))
            for shots in ((4096,) if quick else (1, 4096)):
                add("sample", f"{shots} shot(s), seeded", 1.0, False, _time(st, lambda s=shots: st.sample(11, s), reps))
        norm = st.norm2()
    for r in rows:
        r["peak_gbs"], r["peak_source"] = peak, peak_src
    print(f"norm^2 after the run = {norm:.12f}", file=out)
    return rows


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("n_qubits", type=int, nargs="?", default=30)
    ap.add_argument("--dtype", default="complex128", choices=["complex64", "complex128"])
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--no-observables", action="store_true")
    ap.add_argument("--json", default=None, help="append one JSON line per measurement to this file")
    a = ap.parse_args(argv)
    rows = bench_kernel(a.n_qubits, a.dtype, a.reps, a.device, not a.no_observables)
    if a.json:
        with open(a.json, "a") as f:
            for r in rows:
                f.write(json.dumps(r) + "\n")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
