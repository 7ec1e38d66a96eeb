"""NumPy emulation of the qsv_pass semantics (include/qsv.h) — TEST INFRASTRUCTURE.

Lets the pass compiler be verified on CPU: a compiled Program is executed here on a NumPy
state and compared with the oracle.  It mirrors what pass_kernel.cuh does per tile, including
the descriptor validation rules of qsv.cu (so an invalid plan fails here too)."""
from __future__ import annotations

import numpy as np

from quantum_simulations_b200 import _lib as L
from quantum_simulations_b200.circuit.passes import PassStep, Dense2QStep, Dense1QStep, Program, SwapStep
from oracle import ref_dense as O

R = L.QSV_REG_BITS


def _insert_zero(x, p):
    low = x & ((1 << p) - 1)
    return ((x >> p) << (p + 1)) | low


def run_pass(psi: np.ndarray, desc: L.QsvPass, ops, n_local: int, rank: int = 0, tables=None) -> None:
    t = desc.n_tile
    load = [desc.load_bits[i] for i in range(t)]
    store = [desc.store_bits[i] for i in range(t)]
    assert load == sorted(load) and len(set(load)) == t and all(0 <= b < n_local for b in load)
    assert sorted(store) == load, "store_bits must permute load_bits"
    assert 1 <= desc.n_rounds <= L.QSV_MAX_ROUNDS
    x = np.arange(1 << t, dtype=np.int64)
    off_l = np.zeros_like(x)
    off_s = np.zeros_like(x)
    for i in range(t):
        off_l |= ((x >> i) & 1) << load[i]
        off_s |= ((x >> i) & 1) << store[i]
    base = np.arange(len(psi) >> t, dtype=np.int64)
    for p in load:
        base = _insert_zero(base, p)
    glob = (rank << n_local) | base
    tile_mask = sum(1 << b for b in load)
    if desc.zero_input:
        # fused |0...0> initialisation: the tiles the pass VISITS are not read (treated as |0...0>);
        # tiles it skips (zero-support) keep whatever they hold, so that combination needs a real memset
        live = np.ones(len(base), dtype=bool) if desc.n_active < 0 else \
            (base & ~sum(1 << desc.active_bits[k] for k in range(desc.n_active))) == 0
        psi[(base[live][:, None] + off_l[None, :]).ravel()] = 0
        if rank == 0:
            psi[0] = 1
    if desc.n_active >= 0:
        # zero-support skipping: every tile outside the active set must be exactly zero (the kernel
        # does not visit it); a planner bug shows up here
        act = [desc.active_bits[k] for k in range(desc.n_active)]
        assert act == sorted(act) and all(0 <= b < n_local and not (tile_mask >> b) & 1 for b in act)
        amask = sum(1 << b for b in act)
        dead = (base & ~amask) != 0
        assert not np.any(psi[base[dead][:, None] + off_l[None, :]]), "skipped tile holds data"
    work = psi[base[:, None] + off_l[None, :]]
    for r in range(desc.n_rounds):
        rd = desc.rounds[r]
        regs = [rd.reg_pos[b] for b in range(R)]
        thr = [rd.thr_pos[i] for i in range(t - R)]
        assert sorted(regs + thr) == list(range(t)), f"round {r}: reg/thr positions do not partition the tile"
        assert 0 <= rd.op_begin <= rd.op_end <= desc.n_ops
        regmask = sum(1 << i for i in regs)
        assert rd.fold_off >= -1
        if rd.fold_off >= 0:
            nthr = 1 << (t - R)
            assert tables is not None and rd.fold_off + nthr <= desc.n_fold == len(tables)
            tab = tables[rd.fold_off:rd.fold_off + nthr]
            assert np.abs(np.abs(tab) - 1).max() < 1e-12
            tix = np.zeros(1 << t, dtype=np.int64)       # thread index that owns tile index x
            for k, i in enumerate(thr):
                tix |= ((x >> i) & 1) << k
            work *= tab[tix][None, :]
        for o in range(rd.op_begin, rd.op_end):
            op = ops[o]
            assert not (op.tile_ctrl & regmask) and not (op.tile_ctrl >> t)
            assert not (op.glob_ctrl & tile_mask)
            assert not (op.reg_ctrl >> R)
            rows = (glob & op.glob_ctrl) == op.glob_ctrl
            need = op.tile_ctrl | sum(1 << regs[b] for b in range(R) if (op.reg_ctrl >> b) & 1)
            cols = (x & need) == need
            m = [op.m[k] for k in range(4)]
            ctrl_any = bool(op.reg_ctrl or op.tile_ctrl or op.glob_ctrl)
            if op.kind == L.OP_TPHASE:
                assert not ctrl_any and not op.flags and 0 <= op.target < R and tables is not None
                toff, goff, mask = int(m[0]), int(m[1]), int(m[2])
                nthr = 1 << (t - R)
                tix = np.zeros(1 << t, dtype=np.int64)
                for k, i in enumerate(thr):
                    tix |= ((x >> i) & 1) << k
                fac = np.ones((len(glob), 1 << t), dtype=np.complex128)
                if toff >= 0:
                    assert toff + nthr <= desc.n_fold
                    fac *= tables[toff:toff + nthr][tix][None, :]
                k = 0
                for r_ in range(8):
                    if mask >> r_ & 1:
                        assert goff >= 0 and goff + 256 * (k + 1) <= desc.n_fold
                        fac *= tables[goff + 256 * k + ((glob >> (8 * r_)) & 255)][:, None]
                        k += 1
                tb = 1 << regs[op.target]
                sel = (x & tb) != 0
                work[:, sel] *= fac[:, sel]
                continue
            if op.flags:
                # HAD / ROT with pre-ops: the control fields are parity masks of a sign on the b half
                assert op.kind in (L.OP_HAD, L.OP_ROT) and not op.reg_ctrl
                assert not (op.flags & ~(L.OPF_PRESIGN | L.OPF_PRENEG | L.OPF_PREPHASE))
                assert (op.flags & L.OPF_PRESIGN) or not (op.tile_ctrl or op.glob_ctrl)
                tb = 1 << regs[op.target]
                x1 = x[(x & tb) != 0]
                par_c = np.array([bin(int(v) & op.tile_ctrl).count("1") for v in x1])
                par_r = np.array([bin(int(g) & op.glob_ctrl).count("1") for g in glob])
                neg = 1 if op.flags & L.OPF_PRENEG else 0
                sgn = 1.0 - 2.0 * ((par_r[:, None] + par_c[None, :] + neg) & 1)
                w = work[:, x1] * sgn
                if op.flags & L.OPF_PREPHASE:
                    t_, s_ = m[2], m[3]
                    assert abs(t_) <= 1 + 1e-12 and abs(s_) <= 1 + 1e-12
                    re, im = w.real.copy(), w.imag.copy()
                    re = re - t_ * im
                    im = im + s_ * re
                    re = re - t_ * im
                    w = re + 1j * im
                work[:, x1] = w
                rows = np.ones(len(glob), dtype=bool)
                cols = np.ones(1 << t, dtype=bool)
                ctrl_any = False
            if op.kind in L.OP_WITH_TARGET:
                assert 0 <= op.target < R and not ((op.reg_ctrl >> op.target) & 1)
                tb = 1 << regs[op.target]
                x0 = x[cols & ((x & tb) == 0)]
                a = work[np.ix_(rows, x0)]
                b = work[np.ix_(rows, x0 | tb)]
                if op.kind == L.OP_HAD:
                    assert not ctrl_any, "HAD cannot be controlled"
                    nb = a - b                     # the kernel's two in-place steps
                    na = 2.0 * a - nb
                elif op.kind == L.OP_ROT:
                    t_, s_ = m[0], m[1]
                    assert abs(t_) <= 1 + 1e-12 and abs(s_) <= 1 + 1e-12
                    na = a - t_ * b
                    nb = b + s_ * na
                    na = na - t_ * nb
                elif op.kind == L.OP_XSWAP:
                    na, nb = b, a
                else:                               # YSWAP: (a, b) -> (-i b, i a)
                    na, nb = -1j * b, 1j * a
                work[np.ix_(rows, x0)] = na
                work[np.ix_(rows, x0 | tb)] = nb
            elif op.kind == L.OP_PHASE:
                t_, s_ = m[0], m[1]
                assert abs(t_) <= 1 + 1e-12 and abs(m[2] ** 2 + m[3] ** 2 - 1) < 1e-12 and m[2] >= 0
                assert abs(s_ - m[3]) < 1e-15 and abs(t_ - m[3] / (1 + m[2])) < 1e-14
                xs = x[cols]
                w = work[np.ix_(rows, xs)]
                re, im = w.real.copy(), w.imag.copy()
                re = re - t_ * im                   # rotation of (re, im) by three shears
                im = im + s_ * re
                re = re - t_ * im
                work[np.ix_(rows, xs)] = re + 1j * im
            elif op.kind == L.OP_SIGN:
                xs = x[cols]
                work[np.ix_(rows, xs)] = -work[np.ix_(rows, xs)]
            elif op.kind == L.OP_SCALE:
                assert not ctrl_any
                work *= m[0]
            else:
                raise AssertionError(f"emulator: unsupported op kind {op.kind}")
    assert not (desc.store_flip & ~tile_mask)
    psi[base[:, None] + (off_s ^ int(desc.store_flip))[None, :]] = work


def run_program(prog: Program, psi: np.ndarray, rank: int = 0) -> np.ndarray:
    for step in prog.steps:
        if isinstance(step, PassStep):
            run_pass(psi, step.desc, step.ops, prog.n_local, rank, step.tables)
        elif isinstance(step, Dense2QStep):
            O.apply_2q(psi, step.qa_pos, step.qb_pos, step.U)
        elif isinstance(step, Dense1QStep):
            O.apply_1q(psi, step.q_pos, step.U)
        else:
            raise AssertionError(type(step))
    return psi


def swap_bits_full(psi: np.ndarray, n: int, global_bits, local_bits) -> np.ndarray:
    """SwapStep on the FULL 2^n vector (rank bits = top index bits): exchange index bits."""
    t = psi.reshape((2,) * n)                       # axis k <-> index bit n-1-k
    for g, l in zip(global_bits, local_bits):
        t = np.swapaxes(t, n - 1 - g, n - 1 - l)
    return np.ascontiguousarray(t).reshape(-1)


def run_program_sharded(prog: Program, psi: np.ndarray) -> np.ndarray:
    """Execute a program compiled for n_local < n on all 2^(n - n_local) shards of `psi`."""
    n, n_loc = prog.n_qubits, prog.n_local
    world = 1 << (n - n_loc)
    for step in prog.steps:
        if isinstance(step, PassStep):
            shards = psi.reshape(world, 1 << n_loc)
            for r in range(world):
                run_pass(shards[r], step.desc, step.ops, n_loc, r, step.tables)
        elif isinstance(step, SwapStep):
            assert all(g >= n_loc for g in step.global_bits)
            assert all(0 <= l < n_loc for l in step.local_bits) and len(set(step.local_bits)) == len(step.local_bits)
            psi = swap_bits_full(psi, n, step.global_bits, step.local_bits)
        else:
            raise AssertionError(type(step))
    return psi
