"""Design-space sweep of the pass kernel on a real B200 (run under gpurun).

Writes one JSON line per measurement to gpurun_out/sweep.jsonl:
  * identity passes: achieved GB/s vs tile bits t and contiguous low bits a
  * identity passes with R rounds: the cost of a shared-memory exchange
  * one round with G ops of each kind: the cost of the arithmetic
  * the per-gate kernels for comparison
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from quantum_simulations_b200 import _lib as L                       # noqa: E402
from quantum_simulations_b200.kernel.cuda import DeviceState         # noqa: E402
from quantum_simulations_b200.circuit.passes import PassStep         # noqa: E402
from quantum_simulations_b200.kernel import gates as G               # noqa: E402

OUT = ROOT / "gpurun_out"
OUT.mkdir(exist_ok=True)


def make_pass(n, t, a, rounds=1, ops_per_round=0, kind=L.OP_ROT, high="top", W=3):
    """Identity-layout pass: tile = low a positions + (t-a) high positions."""
    if high == "top":
        hi = list(range(n - (t - a), n))
    else:  # spread
        hi = sorted(set(np.linspace(a, n - 1, t - a).astype(int).tolist()))
        k = n - 1
        while len(hi) < t - a:
            if k not in hi:
                hi.append(k)
            k -= 1
        hi = sorted(hi)
    bits = list(range(a)) + hi
    d = L.QsvPass()
    d.n_tile = t
    for i, b in enumerate(bits):
        d.load_bits[i] = b
        d.store_bits[i] = b
    d.n_rounds = rounds
    ops = []
    for r in range(rounds):
        rd = d.rounds[r]
        if r == 0 or r == rounds - 1:
            regs = list(range(t - 4, t))
        else:
            # rotate register sets through the tile so every exchange really moves data
            start = (4 * r) % (t - 4)
            regs = [(start + j) % t for j in range(4)]
            regs = sorted(set(regs))
            k = t - 1
            while len(regs) < 4:
                if k not in regs:
                    regs.append(k)
                k -= 1
            regs = sorted(regs)
        for b in range(4):
            rd.reg_pos[b] = regs[b]
        free = [i for i in range(t) if i not in regs]
        if 0 < r < rounds - 1:
            head, seen = [], set()
            for i in free:
                if i % W not in seen:
                    head.append(i); seen.add(i % W)
                if len(head) == W:
                    break
            free = head + [i for i in free if i not in head]
        for k_, i in enumerate(free):
            rd.thr_pos[k_] = i
        rd.op_begin = len(ops)
        rd.fold_off = -1
        for g in range(ops_per_round):
            o = L.QsvOp()
            fold = isinstance(kind, str)
            kind_ = {"PHASE_FOLD": L.OP_PHASE, "SIGN_FOLD": L.OP_SIGN}.get(kind, kind)
            o.kind = kind_
            o.target = g % 4
            th = 0.3 + 0.07 * (g % 16)
            if kind_ == L.OP_ROT:
                flat = [np.tan(th / 2), np.sin(th), np.cos(th), 0.0]
            elif kind_ == L.OP_PHASE:
                flat = [np.tan(th / 2), np.sin(th), np.cos(th), np.sin(th)]
            elif kind_ == L.OP_SCALE:
                flat = [1.0000001, 0, 0, 0]
            else:
                flat = [0.0] * 4
            if kind_ in (L.OP_PHASE, L.OP_SIGN):
                if fold:    # controls on two thread-fixed tile positions
                    free_pos = [i for i in range(t) if i not in regs]
                    o.tile_ctrl = (1 << free_pos[g % len(free_pos)]) | (1 << free_pos[(g + 3) % len(free_pos)])
                else:
                    o.reg_ctrl = 1 << (g % 4)
            for k_ in range(4):
                o.m[k_] = float(flat[k_])
            ops.append(o)
        rd.op_end = len(ops)
    d.n_ops = len(ops)
    arr = (L.QsvOp * max(len(ops), 1))(*ops)
    return PassStep(d, arr, len(ops), set(), bits)


def time_pass(st, step, reps=5):
    st.apply_pass(step)
    st.sync()
    st.timing(True)
    for _ in range(reps):
        st.apply_pass(step)
    tm = [ms for ms, kind, _ in st.take_timings() if kind == 10]
    st.timing(False)
    return float(np.median(tm)), float(min(tm))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    dtype = sys.argv[2] if len(sys.argv) > 2 else "complex128"
    ab = np.dtype(dtype).itemsize
    W = 3 if ab == 16 else 4
    tmax = 12 if ab == 16 else 13
    gb = 2 * ab * (1 << n) / 1e9
    rows = []

    def emit(**kw):
        rows.append(kw)
        print(json.dumps(kw), flush=True)

    with DeviceState(n, dtype) as st:
        st.init_zero()
        # 1. identity pass: t x a
        for t in (tmax - 1, tmax):
            for a in (2, 3, 4, 5, 6, 7):
                if a > t - 4:
                    continue
                for high in ("top", "spread"):
                    med, best = time_pass(st, make_pass(n, t, a, 1, 0, high=high, W=W))
                    emit(exp="identity", n=n, dtype=dtype, t=t, a=a, high=high, ms=med, ms_min=best, gbs=gb / med * 1e3)
        # 2. rounds
        for t in (tmax - 1, tmax):
            for R in (1, 2, 3, 4, 5, 6, 8):
                med, best = time_pass(st, make_pass(n, t, 5, R, 0, W=W))
                emit(exp="rounds", n=n, dtype=dtype, t=t, a=5, rounds=R, ms=med, ms_min=best, gbs=gb / med * 1e3)
        # 3. arithmetic in a single round and spread over 3 rounds
        tops = tmax - 1 if ab == 16 else tmax
        for kind, nm in ((L.OP_ROT, "ROT"), (L.OP_HAD, "HAD"), (L.OP_XSWAP, "XSWAP"),
                         (L.OP_YSWAP, "YSWAP"), (L.OP_PHASE, "PHASE"), (L.OP_SIGN, "SIGN"),
                         ("PHASE_FOLD", "PHASE_FOLD"), ("SIGN_FOLD", "SIGN_FOLD")):
            for R in (1, 3):
                for g in (2, 8, 16, 32):
                    med, best = time_pass(st, make_pass(n, tops, 5, R, g, kind=kind, W=W))
                    emit(exp="ops", n=n, dtype=dtype, t=tops, a=5, rounds=R, kind=nm, ops_per_round=g,
                         total_ops=g * R, ms=med, ms_min=best, gbs=gb / med * 1e3)
        # 4. per-gate kernels
        st.timing(True)
        H = G.H()
        for q in (0, 1, 4, 10, 20, n - 1):
            for _ in range(3):
                st.apply_1q(q, H)
            tm = [ms for ms, _, _ in st.take_timings()]
            emit(exp="apply_1q", n=n, dtype=dtype, q=q, ms=float(np.median(tm)), gbs=gb / float(np.median(tm)) * 1e3)
        for qa, qb in ((0, 1), (3, 17), (n - 1, n - 2)):
            for _ in range(3):
                st.apply_2q(qa, qb, G.CNOT() @ np.kron(G.H(), G.T()))
            tm = [ms for ms, _, _ in st.take_timings()]
            emit(exp="apply_2q", n=n, dtype=dtype, qa=qa, qb=qb, ms=float(np.median(tm)), gbs=gb / float(np.median(tm)) * 1e3)
        for qs in ([3], [2, 20]):
            for _ in range(3):
                st.apply_diag(qs, np.exp(1j * np.arange(1 << len(qs))))
            tm = [ms for ms, _, _ in st.take_timings()]
            emit(exp="apply_diag", n=n, dtype=dtype, qs=qs, ms=float(np.median(tm)), gbs=gb / float(np.median(tm)) * 1e3)
        st.timing(False)
    with open(OUT / f"sweep_n{n}_{dtype}.jsonl", "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
