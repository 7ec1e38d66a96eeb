"""Pass compiler: step-IR ops [(phys_qubits, U)] -> fused on-chip passes for libqsv.

This is `batch_levels` (reference wenbo_engine/circuit/fusion.py:86-142) and the Atlas-style
stage selection (reference circuit/staging.py:320-421) moved one level down the memory
hierarchy.  The reference asks "which gates can run while this CHUNK FILE is in RAM?"; here
the question is "which gates can run while this 2^t-amplitude TILE is on chip?" and, inside
a pass, "while these 4 index bits are in a thread's REGISTERS?".  One pass costs exactly one
read + one write of the shard in HBM, whatever it executes, so the compiler minimises passes.

Vocabulary
  content    a qubit's worth of index information.  Initially content c is IR qubit c.  SWAP
             gates never move data: they rename which content an IR qubit refers to.
  position   a bit of the amplitude index in device memory; pos[c] = position of content c.
             A pass may permute the contents of its tile among the tile's positions for free
             (store_bits), which is how layouts are rotated and finally restored.
  micro-op   one kernel op (include/qsv.h QSV_OP_*): an optional TARGET content that is mixed
             (must be a register bit of its round) plus CONTROL contents that are only
             inspected (diagonal action: controls of CNOT/CY/CU and every qubit of
             Z,S,T,R,CZ,CR — Atlas's "insular" qubits, reference staging.py:65-98).  Controls
             may be register bits, other tile bits, bits outside the tile or rank bits.

Ordering rule (this is what the reference's heuristic stager gets wrong):
two micro-ops may be reordered iff on every shared content BOTH act diagonally.  Passes and
rounds always execute a dependency-closed set, so per-qubit program order is preserved.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field, replace

import numpy as np

from quantum_simulations_b200 import _lib as L

REG_BITS = L.QSV_REG_BITS
RING_TILE_BITS = 11
_ONE = 1.0 + 0.0j
_Z8 = (0.0,) * 8


# --------------------------------------------------------------------------- lowering
@dataclass
class MicroOp:
    kind: int
    target: int | None                 # content that is mixed (None: purely diagonal)
    ctrls: tuple[int, ...]             # contents that must be 1 for the op to act
    m: tuple[float, ...] = (0.0,) * 4  # coefficients (see qsv.h)
    src: int = -1                      # IR op index it came from
    seq: int = -1                      # position in the lowered stream (stable order key)
    # pre-ops folded into an uncontrolled HAD / ROT (qsv.h QSV_OPF_*): diagonal action on the b half
    pre_neg: bool = False              # pending Z on the target
    pre_par: frozenset = frozenset()   # partners of the CZ gates pending on the target (parity)
    pre_phase: complex = 1.0 + 0.0j    # pending diag(1, e^{i phi}) on the target
    tph: dict | None = None            # OP_TPHASE: {partner content: unit phase applied when target AND partner are 1}

    looks: tuple = field(init=False, repr=False, compare=False, default=())

    tmask: int = field(init=False, repr=False, compare=False, default=0)    # 1 << target (0: none)
    lmask: int = field(init=False, repr=False, compare=False, default=0)    # OR of 1 << c over `looks`

    def __post_init__(self):
        # contents the op inspects without mixing them (controls + pre-sign partners + table partners)
        out = self.ctrls + tuple(self.pre_par) if self.pre_par else self.ctrls
        self.looks = out + tuple(self.tph) if self.tph else out
        self.tmask = 0 if self.target is None else 1 << self.target
        m = 0
        for c in self.looks:
            m |= 1 << c
        self.lmask = m


@dataclass
class Dense2Q:
    """2-qubit op without diagonal / controlled / swap structure: runs as its own kernel."""
    qa: int
    qb: int
    U: np.ndarray
    src: int = -1


@dataclass
class Dense1Q:
    """1-qubit op that is not unitary (cannot be written with the in-place primitives)."""
    q: int
    U: np.ndarray
    ctrl: int | None = None
    src: int = -1


_Z4 = (0.0,) * 4
_H_EXACT = np.array([[1, 1], [1, -1]], dtype=np.complex128) * (1.0 / np.sqrt(2.0))
_X_EXACT = np.array([[0, 1], [1, 0]], dtype=np.complex128)
_Y_EXACT = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
INV_SQRT2 = 1.0 / np.sqrt(2.0)
_UNITARY_TOL = 1e-13


_SNAP = 4e-16      # phases within a couple of ulp of 1, -1, i, -i are snapped to them


def _phase(d: complex, ctrls: tuple[int, ...], src: int) -> list[MicroOp]:
    """amp *= d (|d| = 1) where all ctrls are 1 -> optional SIGN + PHASE with |phi| <= pi/2."""
    d = complex(d)
    if abs(d.imag) <= _SNAP:
        d = complex(1.0 if d.real > 0 else -1.0, 0.0)
    elif abs(d.real) <= _SNAP:
        d = complex(0.0, 1.0 if d.imag > 0 else -1.0)
    if d == _ONE:
        return []
    out: list[MicroOp] = []
    if d.real < 0:                      # e^{i phi} = -e^{i (phi -+ pi)}
        out.append(MicroOp(L.OP_SIGN, None, ctrls, _Z4, src))
        d = -d
        if d == _ONE:
            return out
    c, sn = d.real, d.imag              # c >= 0 : phi in [-pi/2, pi/2]
    norm = np.hypot(c, sn)
    c, sn = c / norm, sn / norm
    t = sn / (1.0 + c)                  # tan(phi/2), |t| <= 1
    out.append(MicroOp(L.OP_PHASE, None, ctrls, (float(t), float(sn), float(c), float(sn)), src))
    return out


def _rot(c: float, sn: float, q: int, ctrls: tuple[int, ...], src: int) -> list[MicroOp]:
    """(a, b) -> (c a - sn b, sn a + c b) with c >= 0 (|theta| <= pi/2)."""
    if sn == 0.0:
        return []
    t = sn / (1.0 + c)
    return [MicroOp(L.OP_ROT, q, ctrls, (float(t), float(sn), float(c), 0.0), src)]


def lower_1q(u, q: int, ctrls: tuple[int, ...] = (), src: int = -1) -> list:
    """2x2 unitary `u` on content q, active where all `ctrls` are 1, as in-place primitives."""
    u = np.asarray(u, dtype=np.complex128)
    if u[0, 1] == 0 and u[1, 0] == 0:                       # diagonal
        d0, d1 = complex(u[0, 0]), complex(u[1, 1])
        if abs(abs(d0) - 1) > _UNITARY_TOL or abs(abs(d1) - 1) > _UNITARY_TOL:
            return [Dense1Q(q, u, ctrls[0] if ctrls else None, src)] if len(ctrls) <= 1 else _not_lowerable()
        rel = d1 if d0 == _ONE else d1 / d0
        return _phase(d0, ctrls, src) + _phase(rel, ctrls + (q,), src)
    if np.array_equal(u, _X_EXACT):
        return [MicroOp(L.OP_XSWAP, q, ctrls, _Z4, src)]
    if np.array_equal(u, _Y_EXACT):
        return [MicroOp(L.OP_YSWAP, q, ctrls, _Z4, src)]
    if not ctrls and np.array_equal(u, _H_EXACT):
        # unnormalised butterfly; the 1/sqrt2 becomes part of the pass's single SCALE op
        return [MicroOp(L.OP_HAD, q, (), _Z4, src),
                MicroOp(L.OP_SCALE, None, (), (INV_SQRT2, 0.0, 0.0, 0.0), src)]
    if np.abs(u.conj().T @ u - np.eye(2)).max() > _UNITARY_TOL:
        if len(ctrls) <= 1:
            return [Dense1Q(q, u, ctrls[0] if ctrls else None, src)]
        return _not_lowerable()
    if u[0, 0] == 0 and u[1, 1] == 0:                       # anti-diagonal: X then a diagonal
        return [MicroOp(L.OP_XSWAP, q, ctrls, _Z4, src)] + lower_1q(
            np.array([[u[0, 1], 0], [0, u[1, 0]]]), q, ctrls, src)
    # ZYZ: u = e^{ia} diag(1, e^{ip}) [[c, -S], [S, c]] diag(1, e^{il}),  c = |u00| > 0, S real
    a = u[0, 0] / abs(u[0, 0])                              # e^{ia}
    c = float(abs(u[0, 0]))
    w10 = u[1, 0] / a                                       # = e^{ip} S
    w01 = -u[0, 1] / a                                      # = e^{il} S
    S = float(abs(w10))
    ep = w10 / S
    if ep.real < 0:                                         # keep |p| <= pi/2 by signing S
        ep, S = -ep, -S
    el = w01 / S
    norm = np.hypot(c, S)
    ops = _phase(complex(el), ctrls + (q,), src)
    ops += _rot(c / norm, S / norm, q, ctrls, src)
    ops += _phase(complex(ep), ctrls + (q,), src)
    ops += _phase(complex(a), ctrls, src)
    return ops


def _not_lowerable():
    raise NotImplementedError("non-unitary multi-controlled block cannot be lowered to pass ops")


_SWAP = np.eye(4)[[0, 2, 1, 3]]
_I2, _O2 = np.eye(2), np.zeros((2, 2))


def lower_op(qubits, U, src: int = -1) -> list:
    """One step-IR op -> micro-ops | ('swap', a, b) | Dense1Q | Dense2Q.  Structure is detected
    on exact zeros/ones, which gate constructors and products of structured matrices preserve."""
    U = np.asarray(U, dtype=np.complex128)
    if len(qubits) == 1:
        return lower_1q(U, qubits[0], (), src)
    qa, qb = qubits
    if qa == qb:
        raise ValueError("2-qubit op on a repeated qubit")
    if np.array_equal(U, _SWAP):
        return [("swap", qa, qb)]
    if not np.any(U - np.diag(np.diag(U))):
        d = [complex(x) for x in np.diag(U)]                 # index = 2*bit(qa) + bit(qb)
        if max(abs(abs(x) - 1) for x in d) > _UNITARY_TOL:
            return [Dense2Q(qa, qb, U, src)]
        s = d[0]
        pb = d[1] if s == _ONE else d[1] / s
        pa = d[2] if s == _ONE else d[2] / s
        pab = d[3] if (s == _ONE and pa == _ONE and pb == _ONE) else d[3] / (s * pa * pb)
        return (_phase(s, (), src) + _phase(pb, (qb,), src) + _phase(pa, (qa,), src)
                + _phase(pab, (qa, qb), src))
    if np.array_equal(U[:2, :2], _I2) and np.array_equal(U[:2, 2:], _O2) and np.array_equal(U[2:, :2], _O2):
        low = lower_1q(U[2:, 2:], qb, (qa,), src)            # control = qubits[0]
        return [Dense2Q(qa, qb, U, src)] if any(isinstance(x, Dense1Q) for x in low) else low
    ev, od = [0, 2], [1, 3]
    if np.array_equal(U[np.ix_(ev, ev)], _I2) and not np.any(U[np.ix_(ev, od)]) and not np.any(U[np.ix_(od, ev)]):
        low = lower_1q(U[np.ix_(od, od)], qa, (qb,), src)    # control = qubits[1]
        return [Dense2Q(qa, qb, U, src)] if any(isinstance(x, Dense1Q) for x in low) else low
    return [Dense2Q(qa, qb, U, src)]


def _ops_fingerprint(ir_ops) -> bytes:
    """Content hash of a step-IR op list [(qubits, U)]: the key of the per-compiler lowering cache."""
    import hashlib
    h = hashlib.blake2b(digest_size=16)
    h.update(len(ir_ops).to_bytes(8, "little"))
    for qs, U in ir_ops:
        h.update(bytes([len(qs)]))
        h.update(np.asarray(list(qs), dtype=np.int64).tobytes())
        h.update(np.ascontiguousarray(U, dtype=np.complex128).tobytes())
    return h.digest()


_LOWER_CACHE: dict = {}


def lower_op_cached(qubits, U, src: int = -1) -> list:
    """lower_op with the decomposition cached by matrix: circuits repeat a handful of gates (H, X, S,
    T, CZ, ...), and the structure detection + ZYZ of lower_op is most of the lowering time.  The
    cache holds the items for the symbolic qubits (0, 1); they are re-targeted per use."""
    U = np.ascontiguousarray(U, dtype=np.complex128)
    key = (len(qubits), U.tobytes())
    tmpl = _LOWER_CACHE.get(key)
    if tmpl is None:
        if len(_LOWER_CACHE) > 4096:
            _LOWER_CACHE.clear()
        tmpl = lower_op(list(range(len(qubits))), U, -1)
        _LOWER_CACHE[key] = tmpl
    qs = list(qubits)
    if len(qs) == 2 and qs[0] == qs[1]:
        raise ValueError("2-qubit op on a repeated qubit")
    out = []
    for it in tmpl:
        if isinstance(it, tuple):                       # ('swap', 0, 1)
            out.append((it[0], qs[it[1]], qs[it[2]]))
        elif isinstance(it, MicroOp):
            out.append(replace(it, target=None if it.target is None else qs[it.target],
                               ctrls=tuple(qs[c] for c in it.ctrls), src=src))
        elif isinstance(it, Dense2Q):
            out.append(replace(it, qa=qs[it.qa], qb=qs[it.qb], src=src))
        else:                                           # Dense1Q
            out.append(replace(it, q=qs[it.q], ctrl=None if it.ctrl is None else qs[it.ctrl], src=src))
    return out


def absorb_diagonals(rops: list, regset: set) -> tuple[list, int]:
    """Fold the diagonal micro-ops of one round into the mixing op that follows them.

    A diagonal op D commutes with everything up to the first later op that MIXES one of its
    contents.  If that op M is an uncontrolled HAD / ROT on target t and D is
        SIGN{t}        (a Z on the target)                      -> M.pre_neg
        SIGN{t, p}     (a CZ; p not register-resident)          -> M.pre_par ^= {p}
        PHASE{t}       (diag(1, e^{i phi}) on the target)       -> M.pre_phase *= e^{i phi}
    then D is executed by M's record (no dispatch, no divergence: the sign is a per-thread
    xor mask, the phase three more shears on the b half).  Returns (ops, n_absorbed)."""
    out = list(rops)
    n_abs = 0
    for i, d in enumerate(rops):
        if d.target is not None or d.kind not in (L.OP_SIGN, L.OP_PHASE) or not d.ctrls:
            continue
        cs = set(d.ctrls)
        j = next((k for k in range(i + 1, len(rops)) if rops[k].target in cs), None)
        if j is None:
            continue
        m = out[j]
        if m.kind not in (L.OP_HAD, L.OP_ROT) or m.ctrls:
            continue
        t = m.target
        if d.kind == L.OP_SIGN and cs == {t}:
            m = replace(m, pre_neg=not m.pre_neg)
        elif d.kind == L.OP_SIGN and len(cs) == 2 and not ((cs - {t}) & regset):
            m = replace(m, pre_par=m.pre_par ^ frozenset(cs - {t}))
        elif d.kind == L.OP_PHASE and cs == {t}:
            m = replace(m, pre_phase=m.pre_phase * complex(d.m[2], d.m[3]))
        else:
            continue
        out[j] = m
        out[i] = None
        n_abs += 1
    return [o for o in out if o is not None], n_abs


def group_table_phases(rops: list, regset: set) -> tuple[list, int]:
    """Collect the controlled-phase micro-ops of one round that pair a REGISTER content t with a
    content that is not register-resident (PHASE{t, p} / SIGN{t, p}) into one OP_TPHASE per target
    and interval between mixing ops on t.  They are diagonal, so they commute with each other and
    with everything in between that does not MIX t (p cannot be mixed in this round).  A group of
    fewer than three stays as it is (the lifting form of a single phase is cheaper)."""
    pending: dict = {}
    order: dict = {}
    out: list = []

    def flush(t):
        grp = pending.pop(t, None)
        if not grp:
            return
        members = order.pop(t)
        if len(members) < 3:
            out.extend(members)
        else:
            out.append(MicroOp(L.OP_TPHASE, t, (), _Z4, members[0].src, tph=grp))

    n_grouped = 0
    for op in rops:
        if op.target is None and op.kind in (L.OP_PHASE, L.OP_SIGN) and len(op.ctrls) == 2:
            cs = set(op.ctrls)
            in_reg = cs & regset
            if len(in_reg) == 1:
                (t,) = in_reg
                (p_,) = cs - in_reg
                val = -_ONE if op.kind == L.OP_SIGN else complex(op.m[2], op.m[3])
                grp = pending.setdefault(t, {})
                grp[p_] = grp.get(p_, _ONE) * val
                order.setdefault(t, []).append(op)
                n_grouped += 1
                continue
        if op.target is not None:
            flush(op.target)
        out.append(op)
    for t in list(pending):
        flush(t)
    return out, n_grouped


# ------------------------------------------------------------------- Pauli-X frame
def _inverse(op: MicroOp) -> MicroOp:
    if op.kind == L.OP_PHASE:
        t, s, c, s2 = op.m
        return replace(op, m=(-t, -s, c, -s2))
    if op.kind == L.OP_ROT:
        t, s, c, z = op.m
        return replace(op, m=(-t, -s, c, z))
    return op                                   # SIGN, XSWAP, YSWAP are involutions


def _conj_x(U: np.ndarray, which: tuple[bool, ...]) -> np.ndarray:
    """(X^f (x) ...) U (X^f (x) ...) for a dense block; which[i] = flip qubit i (MSB first)."""
    k = len(which)
    mask = sum(1 << (k - 1 - i) for i, f in enumerate(which) if f)
    idx = [r ^ mask for r in range(1 << k)]
    return np.asarray(U)[np.ix_(idx, idx)]


def frame_transform(item, xf: list) -> list:
    """Rewrite one lowered item for a state stored with some index bits flipped.

    xf[c] = 1 means the stored state is X_c |true state>.  X (and the X part of Y) gates are
    never executed: they toggle xf and every later op is conjugated instead —
      control on a flipped content :  C_c(O) -> O(without c) . C_c(O^-1)
      rotation on a flipped target :  X R(theta) X = R(-theta)
      Hadamard on a flipped target :  H X = Z H  (the flip is absorbed)
      Y = i X Z; X Y X = -Y
    The flips still pending at the end are applied by the final passes as an XOR on the store
    address (qsv_pass.store_flip) — no arithmetic, no data movement."""
    if isinstance(item, Dense2Q):
        fa, fb = bool(xf[item.qa]), bool(xf[item.qb])
        return [replace(item, U=_conj_x(item.U, (fa, fb)))] if (fa or fb) else [item]
    if isinstance(item, Dense1Q):
        return [replace(item, U=_conj_x(item.U, (True,)))] if xf[item.q] else [item]
    op: MicroOp = item
    for c in op.ctrls:
        if xf[c]:
            rest = tuple(x for x in op.ctrls if x != c)
            out = frame_transform(replace(op, ctrls=rest), xf)
            xf[c] = 0                                   # treat c as a plain control for the inverse
            try:
                out += frame_transform(_inverse(op), xf)
            finally:
                xf[c] = 1
            return out
    t = op.target
    if t is None:
        return [op]
    if op.kind == L.OP_HAD:
        if xf[t]:
            xf[t] = 0
            return [op, MicroOp(L.OP_SIGN, None, (t,), _Z4, op.src)]
        return [op]
    if op.kind == L.OP_ROT:
        return [_inverse(op)] if xf[t] else [op]
    if op.kind == L.OP_XSWAP:
        if not op.ctrls:
            xf[t] ^= 1
            return []
        return [op]
    if op.kind == L.OP_YSWAP:
        if not op.ctrls:                                # Y = i X Z (Z first); X Y X = -Y
            g = -1j if xf[t] else 1j
            xf[t] ^= 1
            return [MicroOp(L.OP_SIGN, None, (t,), _Z4, op.src)] + _phase(g, (), op.src)
        return ([MicroOp(L.OP_SIGN, None, op.ctrls, _Z4, op.src)] if xf[t] else []) + [op]
    return [op]


# ------------------------------------------------------------------------ compiled form
@dataclass
class PassStep:
    desc: L.QsvPass
    ops: C.Array                      # (QsvOp * max(n,1))
    n_micro_ops: int
    src_ops: set
    tile_contents: list
    tables: np.ndarray | None = None  # complex128 fold tables (desc.n_fold entries)
    n_folded: int = 0                 # diagonal micro-ops absorbed into the tables
    n_absorbed: int = 0               # diagonal micro-ops riding as pre-ops of a HAD / ROT


@dataclass
class Dense2QStep:
    qa_pos: int
    qb_pos: int
    U: np.ndarray
    src_ops: set


@dataclass
class Dense1QStep:
    q_pos: int
    U: np.ndarray
    src_ops: set


@dataclass
class SwapStep:
    """Global<->local qubit swap (qsv_swap_global_local): rank bit global_bits[i] <-> local bit
    local_bits[i] = n_local - len + i.  One all-to-all among groups of 2^len ranks."""
    global_bits: list
    local_bits: list


@dataclass
class Program:
    n_qubits: int
    n_local: int
    dtype: str
    steps: list = field(default_factory=list)
    final_pos: list = field(default_factory=list)   # final_pos[q] = position of IR qubit q
    final_flips: list = field(default_factory=list) # final_flips[q] = 1: qubit q is still stored flipped
    rank_flip_mask: int = 0                          # shard r holds logical shard r ^ rank_flip_mask
    fused_init: bool = False                         # the first pass creates |0...0> itself (zero_input)
    stats: dict = field(default_factory=dict)

    @property
    def passes(self):
        return [s for s in self.steps if isinstance(s, PassStep)]


# ---------------------------------------------------------------- what a specialised kernel can hold
# csrc/jit.cuh (kMaxOps, kMaxCoefs): a pass of 2^11 amplitudes is emitted as straight-line code only if it has at
# most 256 micro-ops and reads at most 448 doubles from its kernel-parameter constant bank; a larger pass runs on
# the interpreting kernels at 0.17-0.36 of the HBM roofline instead of ~0.8 (profiles/r01).  The planner therefore
# keeps every such pass inside both limits (PassCompiler(fit_jit=True), the default): two specialised passes beat
# one interpreted pass.
JIT_MAX_OPS, JIT_MAX_COEFS = 256, 448


def jit_coefs(step: "PassStep") -> int:
    """Doubles the specialised kernel of `step` reads as C.c[i] — the count qsvjit::generate() makes (csrc/jit.cuh):
    ROT 2 (sine and tangent of the three shears), a PREPHASE pre-op 2 more, PHASE 2, SCALE 1."""
    n = 0
    for k in range(step.desc.n_ops):
        op = step.ops[k]
        if op.kind == L.OP_ROT:
            n += 2
        elif op.kind == L.OP_PHASE:
            n += 2
        elif op.kind == L.OP_SCALE:
            n += 1
        if op.kind in (L.OP_HAD, L.OP_ROT) and (op.flags & L.OPF_PREPHASE):
            n += 2
    return n


def fits_jit(step: "PassStep") -> bool:
    """True if the pass is within the limits of a run-time specialised kernel (or is not a 2^11-amplitude pass at
    all: those are executed by the simple interpreting kernel whatever their size)."""
    if step.desc.n_tile != RING_TILE_BITS:
        return True
    return step.desc.n_ops <= JIT_MAX_OPS and jit_coefs(step) <= JIT_MAX_COEFS


# -------------------------------------------------------------------- dependency scan
def _scan(ops, mixable, lookahead: int | None = None):
    """(run, missing): ops that can execute, in order, when exactly the contents in `mixable`
    may be mixed; and the dependency-free ops that only lack their target in `mixable`."""
    bt = 0               # contents an earlier pending op MIXES      (bit masks over contents)
    bc = 0               # contents an earlier pending op INSPECTS
    mix = 0
    for c in mixable:
        mix |= 1 << c
    run: list = []
    missing: list = []
    if lookahead is not None and lookahead < len(ops):
        ops = ops[:lookahead]
    for i, op in enumerate(ops):
        tm = op.tmask
        if not (tm & (bt | bc)) and not (op.lmask & bt):      # dependency-free
            if not tm or tm & mix:
                run.append(i)
                continue
            missing.append(i)
        bt |= tm
        bc |= op.lmask
    return run, missing


class PassCompiler:
    """Greedy tile / round selection (see module docstring)."""

    def __init__(self, n_qubits: int, n_local: int | None = None, dtype: str = "complex128",
                 tile_bits: int | None = None, low_bits: int | None = None,
                 max_rounds: int = 3, restore_layout: bool = True, lookahead: int = 4096,
                 ring: bool | None = None, max_ops: int = 380, x_frame: bool = True,
                 merge_diagonals: bool = True, fold_tables: bool = True,
                 defer_diagonals: bool = False, absorb: bool = True, allow_swaps: bool = True,
                 swap_anywhere: bool = False, rank_flips: bool = False, park_off_last_round: bool = True,
                 table_phases: bool = True, eager_flips: bool = True, low_store_round: bool = True,
                 low_store_bits: int | None = 2, park_reorder: bool = False,
                 warp_local_rounds: bool = False, explore_seed: int | None = None, explore_p: float = 0.4,
                 explore_k: int = 3, fit_jit: bool = True):
        # explore_seed (circuit/sharding.plan's search): the tile of a pass normally follows the pending targets in
        # program order; with a seed, each slot is instead drawn (probability explore_p) among the first explore_k
        # distinct candidates.  Every such plan is as valid as the greedy one (the dependency scan is the same);
        # a few hundred of them sometimes contain one with a pass less.
        self.explore_seed, self.explore_p, self.explore_k = explore_seed, explore_p, explore_k
        # True: no pass of 2^11 amplitudes exceeds what a specialised kernel holds (JIT_MAX_OPS micro-ops,
        # JIT_MAX_COEFS coefficients): compile() re-plans with a smaller per-pass op budget until all of them fit
        self.fit_jit = fit_jit
        self._rng = None
        self.n = n_qubits
        self.n_local = n_qubits if n_local is None else n_local
        self.dtype = np.dtype(dtype).name
        if self.dtype not in ("complex64", "complex128"):
            raise ValueError(f"unsupported dtype {dtype}")
        max_t = 12 if self.dtype == "complex128" else 13
        # default tile: 2^11 complex128 (32 KB) = what the persistent ring kernel stages
        default_t = RING_TILE_BITS        # both dtypes: the tile the ring / specialised kernels stage
        self.W = 3 if self.dtype == "complex128" else 4      # log2(128 B / sizeof(amp))
        self.t = min(tile_bits or default_t, max_t, self.n_local)
        # the ring kernel (pass_ring.cuh) fills shared memory through cp.async in the swizzled
        # layout, so its FIRST round may hold any tile position in registers
        self.ring = (self.dtype == "complex128" and self.t == RING_TILE_BITS) if ring is None else ring
        if self.t < REG_BITS:
            raise ValueError(f"pass kernel needs n_local >= {REG_BITS} (got {self.n_local})")
        # 3 contiguous low positions (128 B runs) already stream at full rate when the other tile
        # bits are spread out (profiles/r01 sweep); fewer forced positions = fewer passes
        a = 3 if low_bits is None else low_bits
        # when the tile does not cover the whole shard, leave room for >= 4 free tile slots
        self.a = self.t if self.t >= self.n_local else max(0, min(a, self.t - REG_BITS))
        self.max_rounds = max(2, min(max_rounds, L.QSV_MAX_ROUNDS - 2))
        self.restore_layout = restore_layout
        self.lookahead = lookahead
        self.max_ops = max_ops          # ops of one pass live in shared memory (kRingMaxOps = 400)
        self.x_frame = x_frame
        self.merge_diagonals = merge_diagonals
        self.fold_tables = fold_tables
        self.defer_diagonals = defer_diagonals
        self.absorb = absorb
        self.allow_swaps = allow_swaps and self.n_local < self.n
        # True: a SwapStep may name any local positions >= 5 (peer-memory swap kernel); False: the
        # outgoing qubits are first relabelled onto the top local positions (contiguous blocks)
        self.swap_anywhere = swap_anywhere
        # True: an X still pending on a rank bit at the end is NOT executed; it is reported in
        # Program.rank_flip_mask and means "shard r holds logical shard r ^ mask" (a free renaming
        # of the ranks by the runner).  False: such a qubit is swapped in and its flip materialised.
        self.rank_flips = rank_flips
        self.park_off_last_round = park_off_last_round
        # True: a last round whose registers hold a content that is stored to one of the low (128-byte
        # row) positions is followed by an idle round that moves the registers elsewhere, so that the
        # lanes cover whole rows.  False (experiment): store as is — half-row stores that L2 merges,
        # against one shared-memory round trip less.
        self.low_store_round = low_store_round
        # the idle round before the stores is added only if the last round holds a store position BELOW this in
        # registers.  Default 2 (positions 0 and 1: a thread would otherwise write 16- or 32-byte pieces of a row);
        # position 2 in registers means 64-byte pieces, which cost nothing measurable, while the saved round does
        # (profiles/r02/bench_ab_signs_deferred_lsb2.json: 41.9 -> 40.7 ms).  None = any of the W row positions.
        self.low_store_bits = low_store_bits
        # experiment (off): parked qubits / the padding of the last compute round avoid low store positions, which
        # removes the idle store round of more passes — measured on the headline plan: one pass -0.45 ms, another
        # +0.84 ms (profiles/r02/bench_ab_park_reorder.json), so it is not the default
        self.park_reorder = park_reorder
        # experiment: keep the two tile positions that select the WARP inside a consumer group (thread
        # bits 5, 6) the same from one round to the next wherever both rounds leave them out of the
        # registers.  The shared-memory exchange between such rounds stays inside each warp, and the
        # specialised kernel (QSV_JIT_WARP_SYNC=1) replaces the group's named barrier by __syncwarp().
        self.warp_local_rounds = warp_local_rounds
        self.table_phases = table_phases
        # materialise a pending X as soon as its qubit will never be MIXED again (instead of waiting
        # until nothing inspects it either): the not-yet-scheduled ops that inspect it are re-conjugated
        self.eager_flips = eager_flips

    # ---- public -----------------------------------------------------------------------
    def compile(self, ir_ops, init_pos=None, init_flips=None, home_pos=None, zero_state: bool = False,
                fuse_init: bool = False) -> Program:
        """_compile_once, re-planned with a smaller per-pass op budget while a pass of 2^11 amplitudes exceeds what a
        run-time specialised kernel holds (fit_jit; jit_coefs / fits_jit above).  The budgets tried are fixed, so
        every rank of a sharded run arrives at the same plan; circuits whose passes fit (all BASELINE workloads) are
        planned exactly once, as before."""
        prog = self._compile_once(ir_ops, init_pos, init_flips, home_pos, zero_state, fuse_init)
        if not self.fit_jit or all(fits_jit(s) for s in prog.passes):
            return prog
        keep = self.max_ops
        try:
            for cap in (224, 160, 112, 80, 56):
                if cap >= keep:
                    continue
                self.max_ops = cap
                cand = self._compile_once(ir_ops, init_pos, init_flips, home_pos, zero_state, fuse_init)
                cand.stats["fit_jit_max_ops"] = cap
                prog = cand
                if all(fits_jit(s) for s in prog.passes):
                    break
        finally:
            self.max_ops = keep
        prog.stats["passes_beyond_jit"] = sum(not fits_jit(s) for s in prog.passes)
        return prog

    def _compile_once(self, ir_ops, init_pos=None, init_flips=None, home_pos=None, zero_state: bool = False,
                      fuse_init: bool = False) -> Program:
        """init_pos[q]: physical position of IR qubit q before the first op (default q);
        home_pos[q]: position it must have after the last one (default: init_pos[q]).
        zero_state=True: the caller guarantees the state is |0...0> when the program starts (rank 0
        holds amp[0] = 1): passes then visit only the tiles that can hold data (qsv_pass.n_active) —
        index bits of qubits no pass has had in its tile yet are still 0 everywhere.
        fuse_init=True: the program STARTS from |0...0> whatever the shard holds: its first pass does not
        read its input (qsv_pass.zero_input), so neither a memset nor the read half of that pass is paid."""
        import random
        self._rng = None if self.explore_seed is None else random.Random(self.explore_seed)   # same plan for the same seed
        # positions whose index bit may be 1 somewhere in the stored state; None = all (no skipping)
        self._support = set() if (zero_state and self.n_local == self.n) else None
        n = self.n
        # the lowering depends only on the op list and the initial frame: planning the same circuit
        # for several initial placements (sharding.plan / plan_single) lowers it once
        # (keyed by CONTENT: an id() can be recycled by a new list, and a list can be mutated in place)
        lkey = (_ops_fingerprint(ir_ops), tuple(init_flips) if init_flips is not None else None)
        cached = getattr(self, "_lowered", None)
        alias = list(range(n))                      # IR qubit -> content
        xf = list(init_flips) if init_flips is not None else [0] * n   # Pauli-X frame per content
        segments: list = [[]]
        seq = 0
        if cached is not None and cached[0] == lkey:
            segments, alias, xf = cached[1], list(cached[2]), list(cached[3])
            ir_ops = ()                              # skip the lowering loop below
        # Diagonal micro-ops commute with each other and with everything that only INSPECTS
        # their contents, so they are pooled per control set (phases multiply, signs cancel)
        # and only emitted right before an op that MIXES one of their contents.
        pool: dict = {}

        def emit(item) -> None:
            nonlocal seq
            if isinstance(item, (Dense2Q, Dense1Q)):
                segments.append(item)
                segments.append([])
            else:
                item.seq = seq
                seq += 1
                segments[-1].append(item)

        def flush(touching=None, src=-1) -> None:
            for key in [k for k in pool if touching is None or (k & touching)]:
                for op in _phase(pool.pop(key), tuple(sorted(key)), src):
                    emit(op)

        for i, (qs, U) in enumerate(ir_ops):
            for low in lower_op_cached([alias[q] for q in qs], U, i):
                if isinstance(low, tuple):          # swap: rename, no data movement
                    qa, qb = list(qs)
                    alias[qa], alias[qb] = alias[qb], alias[qa]
                    continue
                for item in (frame_transform(low, xf) if self.x_frame else [low]):
                    if isinstance(item, Dense2Q):
                        flush({item.qa, item.qb}, item.src)
                    elif isinstance(item, Dense1Q):
                        flush({item.q}, item.src)
                    elif item.target is None and item.kind in (L.OP_PHASE, L.OP_SIGN) and self.merge_diagonals:
                        key = frozenset(item.ctrls)
                        val = -1.0 + 0j if item.kind == L.OP_SIGN else complex(item.m[2], item.m[3])
                        pool[key] = pool.get(key, _ONE) * val
                        continue
                    elif item.target is not None:
                        if self.absorb and item.kind in (L.OP_HAD, L.OP_ROT) and not item.ctrls:
                            item = self._absorb_pool(pool, item)
                        flush({item.target}, item.src)
                    emit(item)
        if cached is None or cached[0] != lkey:
            flush()
            self._lowered = (lkey, segments, list(alias), list(xf))
        pos = list(init_pos) if init_pos is not None else list(range(n))
        home = [0] * n                              # home[content] = position it must end at
        for q in range(n):
            home[alias[q]] = home_pos[q] if home_pos is not None else (q if init_pos is None else init_pos[q])
        prog = Program(n, self.n_local, self.dtype)
        live = [s for s in segments if isinstance(s, (Dense2Q, Dense1Q)) or s]
        # uses[c] = micro-ops not yet scheduled that touch content c.  A content with no
        # remaining use can go home / be un-flipped by whichever pass holds it last.
        self._uses = [0] * n
        self._xf, self._home = xf, home
        for seg in live:
            if isinstance(seg, Dense2Q):
                self._uses[seg.qa] += 1
                self._uses[seg.qb] += 1
            elif isinstance(seg, Dense1Q):
                self._uses[seg.q] += 1
            else:
                for op in seg:
                    for c in op.looks + ((op.target,) if op.target is not None else ()):
                        self._uses[c] += 1
        for k, seg in enumerate(live):
            if isinstance(seg, Dense1Q):
                if pos[seg.q] >= self.n_local:
                    raise NotImplementedError(
                        f"non-local gate: content {seg.q} sits on rank bit {pos[seg.q]}")
                if xf[seg.q]:
                    raise AssertionError("frame_transform leaves no flip on a dense block")
                self._uses[seg.q] -= 1
                prog.steps.append(Dense1QStep(pos[seg.q], seg.U, {seg.src}))
            elif isinstance(seg, Dense2Q):
                self._uses[seg.qa] -= 1
                self._uses[seg.qb] -= 1
                for c in (seg.qa, seg.qb):
                    if pos[c] >= self.n_local:
                        raise NotImplementedError(
                            f"non-local gate: content {c} sits on rank bit {pos[c]}")
                prog.steps.append(Dense2QStep(pos[seg.qa], pos[seg.qb], seg.U, {seg.src}))
            else:
                self._plan_segment(seg, pos, home, prog, xf, last_segment=(k == len(live) - 1))
        if self.restore_layout:
            self._restore(prog, pos, home, xf)
        if fuse_init and prog.steps and isinstance(prog.steps[0], PassStep) and self._support is None:
            prog.steps[0].desc.zero_input = 1        # (zero-support skipping needs a really zeroed shard)
            prog.fused_init = True
        prog.final_pos = [pos[alias[q]] for q in range(n)]
        prog.final_flips = [xf[alias[q]] for q in range(n)]
        prog.rank_flip_mask = sum(1 << (prog.final_pos[q] - self.n_local) for q in range(n)
                                  if prog.final_flips[q] and prog.final_pos[q] >= self.n_local)
        ps = prog.passes
        prog.stats = {
            "passes": len(ps), "dense2q_steps": sum(isinstance(x, Dense2QStep) for x in prog.steps),
            "dense1q_steps": sum(isinstance(x, Dense1QStep) for x in prog.steps),
            "micro_ops": sum(s.n_micro_ops for s in ps), "folded_ops": sum(s.n_folded for s in ps),
            "absorbed_ops": sum(s.n_absorbed for s in ps),
            "rounds": sum(s.desc.n_rounds for s in ps),
            "swaps": sum(isinstance(x, SwapStep) for x in prog.steps),
            "swap_bits": sum(len(x.global_bits) for x in prog.steps if isinstance(x, SwapStep)),
            "max_rounds_in_pass": max((s.desc.n_rounds for s in ps), default=0),
            "tile_bits": self.t, "low_bits": self.a,
        }
        return prog

    @staticmethod
    def _absorb_pool(pool: dict, m: MicroOp) -> MicroOp:
        """Pending diagonals on m's target that an uncontrolled HAD / ROT can carry as pre-ops:
        the 1-qubit phase {t} and every pure CZ sign {t, p}.  They leave the pool, so they are
        scheduled WITH the mixing op instead of whenever their contents happen to be on chip."""
        t = m.target
        neg, par, ph = m.pre_neg, set(m.pre_par), m.pre_phase
        for key in [k for k in pool if t in k]:
            val = pool[key]
            if len(key) == 1:
                ph *= val
            elif len(key) == 2 and val == -_ONE:
                par ^= set(key - {t})
            elif val == _ONE:
                pass
            else:
                continue
            del pool[key]
        return replace(m, pre_neg=neg, pre_par=frozenset(par), pre_phase=ph)

    # ---- pass planning ----------------------------------------------------------------
    def _plan_segment(self, ops, pos, home, prog, xf, last_segment):
        remaining = list(ops)
        while remaining:
            tile = self._choose_tile(remaining, pos, forced=self._low_contents(pos))
            run, _ = _scan(remaining, set(tile), self.lookahead)
            if not run:
                _, missing = _scan(remaining, set(tile), self.lookahead)
                stuck = [remaining[i].target for i in missing if pos[remaining[i].target] >= self.n_local]
                if stuck and self.allow_swaps:
                    self._swap_in(remaining, pos, home, prog, xf)
                    continue
                for c in stuck:
                    raise NotImplementedError(
                        f"non-local gate: content {c} sits on rank bit {pos[c]} and must be "
                        "mixed; swap it with a local bit first")
                raise RuntimeError("planner made no progress")
            chosen = [remaining[i] for i in run]
            rounds, deferred = self._plan_rounds(chosen, tile)
            done = {id(op) for r in rounds for op in r[1]}
            for r in rounds:
                for op in r[1]:
                    for c in op.looks + ((op.target,) if op.target is not None else ()):
                        self._uses[c] -= 1
            remaining = [op for op in remaining if id(op) not in done]
            eager: set = set()
            if self.eager_flips and self.x_frame and self.restore_layout and last_segment and remaining:
                mixed_later = {op.target for op in remaining if op.target is not None}
                # ... and is only inspected by DIAGONAL ops from here on (a flipped control of a mixing
                # op would turn into an X on its target, i.e. a frame change of another content)
                ctrl_of_mixing = {c for op in remaining if op.target is not None for c in op.ctrls}
                eager = {c for c in tile if xf[c] and self._uses[c] > 0 and c not in mixed_later
                         and c not in ctrl_of_mixing}
                if eager:
                    remaining = self._reconjugate(remaining, eager)
            final = not remaining and last_segment and self.restore_layout
            park = []
            if remaining and self.a:
                wish = self._choose_tile(remaining, pos, forced=[], pool=set(tile))
                # contents whose home is a rank bit wait on the top local slots for the swap that takes
                # them there: they are never parked (a swap of untouched top bits overlaps with the pass)
                cand = [c for c in wish if c in set(tile) and home[c] < self.n_local]
                if self.park_off_last_round and rounds:
                    # a content that is register-resident in the LAST round cannot be stored to a low
                    # position without an extra idle round (the lanes must cover the low positions for
                    # coalesced stores): prefer parking the others
                    last_regs = set(rounds[-1][0])
                    cand = [c for c in cand if c not in last_regs] + [c for c in cand if c in last_regs]
                park = cand[: self.a]
            prog.steps.append(self._emit_pass(tile, rounds, pos, home, park, final, xf, eager=eager))

    def _reconjugate(self, remaining: list, flipped: set) -> list:
        """The stored bit of every content in `flipped` is about to be inverted (its pending X is
        materialised by this pass's store).  Rewrite the unscheduled ops that INSPECT such a content
        so that they act the same on the new storage: a pre-sign partner toggles the constant sign;
        an AND-control c turns C_c(O) into O . C_c(O^-1) (the rule of frame_transform).  None of the
        contents is mixed again (caller's condition), so targets are unaffected."""
        out: list = []
        for op in remaining:
            items = [op]
            for c in flipped:
                nxt = []
                for it in items:
                    if c in it.pre_par:
                        it = replace(it, pre_neg=not it.pre_neg)
                    if c in it.ctrls:
                        one = [0] * self.n
                        one[c] = 1
                        nxt += frame_transform(it, one)
                    else:
                        nxt.append(it)
                items = nxt
            out += items
        # uses[] counts the unscheduled ops per content: recount (an op may have become two)
        self._uses = [0] * self.n
        for op in out:
            for c in op.looks + ((op.target,) if op.target is not None else ()):
                self._uses[c] += 1
        return out

    def _low_contents(self, pos):
        at = {p: c for c, p in enumerate(pos)}
        return [at[p] for p in range(self.a)]

    def _choose_tile(self, remaining, pos, forced, pool=None):
        """t contents for a pass: follow the pending targets in program order.  `forced`
        contents (those sitting in the always-in-tile low positions) are in from the start.
        With `pool` (peeking at the pass after next), at most t-a contents may come from
        outside the pool, because `a` slots of that pass will hold parked pool members."""
        tile = list(forced)
        in_tile = set(tile)
        outside_cap = self.t - self.a if pool is not None else self.t
        n_outside = 0
        rng = self._rng
        while len(tile) < self.t:
            _, missing = _scan(remaining, in_tile, self.lookahead)
            pick = None
            cands: list = []
            for i in missing:
                c = remaining[i].target
                if pos[c] >= self.n_local:
                    continue                       # rank bit: cannot be mixed in this stage
                if pool is not None and c not in pool and n_outside >= outside_cap:
                    continue
                if rng is None:
                    pick = c
                    break
                if c not in cands:
                    cands.append(c)
                    if len(cands) >= self.explore_k:
                        break
            if rng is not None and cands:
                pick = cands[0] if rng.random() >= self.explore_p else rng.choice(cands)
            if pick is None:
                break
            if pool is not None and pick not in pool:
                n_outside += 1
            tile.append(pick)
            in_tile.add(pick)
        if pool is not None:
            return tile
        # pad with idle local contents: first those that still need a free fix-up (pending X
        # flip or displaced from home) and have no remaining use, then highest positions
        if len(tile) < self.t:
            xf, home, uses = self._xf, self._home, self._uses

            def pad_key(c):
                needs_fix = (xf[c] or pos[c] != home[c]) and uses[c] == 0 and self.restore_layout
                # idle slots go to the LOWEST free positions: they extend the contiguous run of a tile
                # row (128 B with 3 low bits, 256 B with 4, ...), which is what DRAM page locality
                # needs when the working qubits are all high (measured: 8.2 -> 5.8 ms per pass)
                return (0 if needs_fix else 1, pos[c])

            for c in sorted((c for c in range(self.n) if pos[c] < self.n_local), key=pad_key):
                if c not in in_tile:
                    tile.append(c)
                    in_tile.add(c)
                    if len(tile) >= self.t:
                        break
        return tile

    def _plan_rounds(self, chosen, tile):
        """Split a pass's ops into rounds of <= 4 register contents.  Returns (rounds, deferred)
        with rounds = [(reg_contents, ops)]; ops beyond max_rounds are deferred to a later pass."""
        pend = list(chosen)
        rounds = []
        total = 0
        while pend and len(rounds) < self.max_rounds and total < self.max_ops:
            regs: list = []
            while len(regs) < REG_BITS:
                _, missing = _scan(pend, set(regs))
                if not missing:
                    break
                regs.append(pend[missing[0]].target)
            run, _ = _scan(pend, set(regs))
            run = run[: self.max_ops - total]          # a prefix is still dependency-closed
            if self.fold_tables and self.defer_diagonals and len(rounds) + 1 < self.max_rounds:
                # A diagonal op with a control in this round's registers costs a register op;
                # if nothing later in this round needs it done, let it wait for a round where
                # all its controls are thread-fixed: there it folds into the round's table.
                regset = set(regs)
                mixed_later: set = set()
                keep = []
                for i in reversed(run):
                    op = pend[i]
                    if (op.target is None and op.ctrls and (regset & set(op.ctrls))
                            and not (mixed_later & set(op.ctrls))):
                        continue                        # deferred
                    if op.target is not None:
                        mixed_later.add(op.target)
                    keep.append(i)
                if any(pend[i].target is not None for i in keep):
                    run = keep[::-1]
            if not run:
                break
            total += len(run)
            rs = set(run)
            rounds.append((regs, [pend[i] for i in run]))
            pend = [op for i, op in enumerate(pend) if i not in rs]
        return rounds, pend

    # ---- pass emission ----------------------------------------------------------------
    def _emit_pass(self, tile, rounds, pos, home, park, final, xf=None, explicit_store=None, eager=frozenset()) -> PassStep:
        t, W = self.t, min(self.W, self.t - REG_BITS)
        load_bits = sorted(pos[c] for c in tile)
        at = {p: c for c, p in enumerate(pos)}
        content = [at[p] for p in load_bits]              # content at tile index i
        idx_of = {c: i for i, c in enumerate(content)}
        finished = {c for c in content if self._uses[c] == 0} if self.restore_layout else set()
        if explicit_store is not None:
            store = [explicit_store[c] for c in content]
        else:
            store = self._choose_store(content, load_bits, home, park, final, finished)   # per tile index
            if rounds and not final and self.park_reorder and self.low_store_bits is not None:
                # Which parked content sits on which of the low (row) positions is free: the next pass loads all of
                # them.  Give the LOWEST positions to contents that are NOT in the registers of the last compute
                # round, so that this pass can store straight from that round (no idle store round: a register-held
                # content on position 0 / 1 would make every thread write 16- / 32-byte pieces of a row).
                last_regs = set(rounds[-1][0])
                movable = [i for i in range(t) if store[i] < W and not (content[i] in finished and store[i] == home[content[i]])]
                slots = sorted(store[i] for i in movable)
                for i, p_ in zip(sorted(movable, key=lambda i: (content[i] in last_regs, store[i])), slots):
                    store[i] = p_

        lo_load = {i for i in range(t) if load_bits[i] < W}
        lo_store = {i for i in range(t) if store[i] < W}
        lo_force = lo_store if self.low_store_bits is None else {i for i in range(t) if store[i] < self.low_store_bits}
        lo_force_pos = {store[i] for i in lo_force}

        # register tile-indices of every round, padded to 4 with idle indices
        plan = []
        for k_, (regs_c, rops) in enumerate(rounds):
            regs = [idx_of[c] for c in regs_c]
            idle = list(range(t - 1, -1, -1))
            if k_ == len(rounds) - 1 and self.park_reorder:
                # the last compute round stores from its registers: pad it with indices that are NOT stored to a
                # low (row) position, or the padding alone would force the idle store round
                idle.sort(key=lambda i: store[i] in lo_force_pos)
            for i in idle:
                if len(regs) >= REG_BITS:
                    break
                if i not in regs:
                    regs.append(i)
            plan.append((regs, rops))
        if not plan:
            plan.append((self._idle_regs(lo_store if self.ring else lo_load | lo_store), []))
        if set(plan[0][0]) & lo_load and not self.ring:
            plan.insert(0, (self._idle_regs(lo_load), []))
        if set(plan[-1][0]) & lo_force and self.low_store_round:
            plan.append((self._idle_regs(lo_store), []))
        if len(plan) == 1:
            # one round = same thread mapping for load and store: only coalesced on both
            # sides if the W lowest positions hold the same tile indices before and after
            free = [i for i in range(t) if i not in plan[0][0]]
            by_load = sorted(free, key=lambda i: load_bits[i])[:W]
            by_store = sorted(free, key=lambda i: store[i])[:W]
            if by_load != by_store and not self.ring:
                plan.append((self._idle_regs(lo_store), []))
        if len(plan) > L.QSV_MAX_ROUNDS:
            raise RuntimeError("too many rounds in one pass")

        desc = L.QsvPass()
        desc.n_tile = t
        for i in range(t):
            desc.load_bits[i] = load_bits[i]
            desc.store_bits[i] = store[i]
        desc.n_rounds = len(plan)
        flat: list = []
        srcs: set = set()
        tables: list = []
        n_folded = 0
        n_absorbed = 0
        g_scale = 1.0                 # product of SCALE micro-ops (1/sqrt2 per Hadamard)
        g_phase = _ONE                # product of uncontrolled PHASE / SIGN micro-ops
        def thread_order(r, regs):
            free = [i for i in range(t) if i not in regs]
            if r == len(plan) - 1:
                return sorted(free, key=lambda i: store[i])
            if r == 0 and not self.ring:
                return sorted(free, key=lambda i: load_bits[i])
            return self._bank_friendly(free)

        thr_of = [thread_order(r, regs) for r, (regs, _) in enumerate(plan)]
        if self.warp_local_rounds and t - REG_BITS == 7:
            # thread bits 0..2 are fixed by coalescing / bank groups; any two of the other four thread
            # positions may sit on bits 5, 6.  Walk the round boundaries backwards and give both rounds a
            # common pair there whenever one exists.
            # (middle rounds of the specialised kernels may use ANY thread position: their bank-group codes
            # are chosen per pass for whatever sits on thread bits 0..2)
            pinned = [False] * len(plan)
            last = len(plan) - 1

            def movable(k):
                return thr_of[k][3:7] if k == last else thr_of[k]

            for r in range(last - 1, -1, -1):
                if r == 0 and not self.ring:
                    break
                nxt = thr_of[r + 1]
                pool = movable(r + 1)
                options = [tuple(nxt[5:7])] if pinned[r + 1] else [(a, b) for a in pool for b in pool if a < b]
                mine = set(movable(r))
                pair = next((p_ for p_ in options if p_[0] in mine and p_[1] in mine), None)
                if pair is None:
                    continue
                for k in (r, r + 1):
                    thr_of[k] = [i for i in thr_of[k] if i not in pair] + list(pair)
                pinned[r] = pinned[r + 1] = True
        for r, (regs, rops) in enumerate(plan):
            rd = desc.rounds[r]
            slot_of = {}
            for b, i in enumerate(sorted(regs)):
                rd.reg_pos[b] = i
                slot_of[content[i]] = b
            thr = thr_of[r]
            for k, i in enumerate(thr):
                rd.thr_pos[k] = i
            rd.op_begin = len(flat)
            rd.fold_off = -1
            fold = None                                      # per-thread phase of this round
            if self.absorb:
                regset = {content[i] for i in regs}
                split = []
                for op in rops:                              # a pre-sign partner that sits in a register
                    inreg = op.pre_par & regset              # slot of this round runs as its own CZ op
                    if inreg:
                        split += [MicroOp(L.OP_SIGN, None, (op.target, c), _Z4, op.src) for c in sorted(inreg)]
                        op = replace(op, pre_par=op.pre_par - inreg)
                    n_absorbed += len(op.pre_par) + int(op.pre_neg) + int(op.pre_phase != _ONE)
                    split.append(op)
                rops, na = absorb_diagonals(split, regset)
                n_absorbed += na
            if self.table_phases:
                rops, _ = group_table_phases(rops, {content[i] for i in regs})
            for op in rops:
                srcs.add(op.src)
                if op.kind == L.OP_TPHASE:
                    flat.append(self._encode_tphase(op, slot_of, idx_of, pos, thr, tables))
                    continue
                if op.kind == L.OP_SCALE:                    # global scalars are not executed
                    g_scale *= op.m[0]                       # where they occur: they commute
                    continue                                 # with everything
                if not op.ctrls and op.kind == L.OP_SIGN:
                    g_phase = -g_phase
                    continue
                if not op.ctrls and op.kind == L.OP_PHASE:
                    g_phase *= complex(op.m[2], op.m[3])
                    continue
                o = self._encode(op, slot_of, idx_of, pos)
                if self.fold_tables and o.kind in (L.OP_PHASE, L.OP_SIGN) and not o.reg_ctrl and not o.glob_ctrl:
                    # controls are thread-fixed tile bits only: commutes with every register op of
                    # the round -> multiply into the round's per-thread table instead of an op
                    if fold is None:
                        fold = np.ones(1 << (t - REG_BITS), dtype=np.complex128)
                        tix = np.arange(1 << (t - REG_BITS))
                        xb_of = np.zeros_like(tix)
                        for k, i in enumerate(thr):
                            xb_of |= ((tix >> k) & 1) << i
                    hit = (xb_of & o.tile_ctrl) == o.tile_ctrl
                    fold[hit] *= -1.0 if o.kind == L.OP_SIGN else complex(o.m[2], o.m[3])
                    n_folded += 1
                    continue
                flat.append(o)
            if fold is not None:
                fold /= np.abs(fold)
                rd.fold_off = sum(len(x) for x in tables)
                tables.append(fold)
            if r == len(plan) - 1:
                if g_phase != _ONE:
                    g_phase /= abs(g_phase)
                    for gop in _phase(complex(g_phase), (), -1):
                        flat.append(self._encode(gop, slot_of, idx_of, pos))
                if g_scale != 1.0:
                    sc = L.QsvOp()
                    sc.kind = L.OP_SCALE
                    sc.m[0] = g_scale
                    flat.append(sc)
            rd.op_end = len(flat)
        desc.n_ops = len(flat)
        tab = np.ascontiguousarray(np.concatenate(tables)) if tables else None
        desc.n_fold = 0 if tab is None else len(tab)
        arr = (L.QsvOp * max(len(flat), 1))(*flat)
        flip = 0
        if xf is not None:                                # materialise pending X gates for free:
            for i in range(t):                            # safe once nothing later looks at the bit
                if xf[content[i]] and (final or content[i] in finished or content[i] in eager):
                    flip |= 1 << store[i]
                    xf[content[i]] = 0
        desc.store_flip = flip
        desc.n_active = -1
        support = getattr(self, "_support", None)
        if support is not None:
            active = sorted(support - set(load_bits))
            if len(active) <= L.QSV_MAX_ACTIVE_BITS and len(active) < self.n_local - t:
                desc.n_active = len(active)
                for k, b in enumerate(active):
                    desc.active_bits[k] = b
            support |= set(load_bits)                     # a pass may leave data on all its tile positions
        for i in range(t):                                # commit the relabelling
            pos[content[i]] = store[i]
        return PassStep(desc, arr, len(flat), srcs, list(tile), tab, n_folded, n_absorbed)

    def _idle_regs(self, avoid) -> list:
        regs = [i for i in range(self.t - 1, -1, -1) if i not in avoid][:REG_BITS]
        if len(regs) < REG_BITS:
            regs = list(range(self.t - REG_BITS, self.t))
        return regs

    def _bank_friendly(self, free) -> list:
        """Middle rounds: the first W thread bits should drive tile indices with distinct
        residues mod W, so a quarter-warp's 128-bit shared accesses hit distinct bank groups
        under tile_swizzle (pass_kernel.cuh)."""
        W = self.W
        free = sorted(free)
        head, seen = [], set()
        for i in free:
            if i % W not in seen:
                head.append(i)
                seen.add(i % W)
            if len(head) == W:
                break
        return head + [i for i in free if i not in head]

    def _choose_store(self, content, load_bits, home, park, final, finished=frozenset()) -> list:
        """New position of the content at every tile index (a permutation of load_bits)."""
        import itertools
        t = len(content)
        cur = {content[i]: load_bits[i] for i in range(t)}
        idx = {content[i]: i for i in range(t)}
        occ = {load_bits[i]: content[i] for i in range(t)}
        new = dict(cur)
        if final:
            slots = set(load_bits)
            new = {}
            for c in content:                             # everyone whose home is here goes home
                if home[c] in slots:
                    new[c] = home[c]
            used = set(new.values())
            for c in content:                             # others keep their slot if still free
                if c not in new and cur[c] not in used:
                    new[c] = cur[c]
                    used.add(cur[c])
            spare = sorted(slots - used)
            for c in content:
                if c not in new:
                    new[c] = spare.pop(0)
            return [new[c] for c in content]
        low = [p for p in load_bits if p < self.a]
        park = [c for c in park if c in cur][: len(low)]
        parked = set(park)
        movers = [c for c in park if cur[c] not in low]
        vacate = [p for p in low if occ[p] not in parked][: len(movers)]
        movers = movers[: len(vacate)]
        if not movers:
            sent = self._send_home(new, content, home, finished, parked)      # ONE call: it mutates
            return [sent[c] for c in content]

        def assign(order):
            out = dict(cur)
            for c, p in zip(order, vacate):
                out[c], out[occ[p]] = p, cur[c]
            return out

        def score(asg):                                   # distinct bank residues at positions < W
            byp = {p: c for c, p in asg.items()}
            res = {idx[byp[p]] % self.W for p in range(min(self.W, self.a)) if p in byp}
            return len(res)

        best, best_s = None, -1
        for order in itertools.islice(itertools.permutations(movers), 120):
            asg = assign(order)
            sc = score(asg)
            if sc > best_s:
                best, best_s = asg, sc
        sent = self._send_home(best, content, home, finished, parked)
        return [sent[c] for c in content]

    def _send_home(self, asg: dict, content, home, finished, parked) -> dict:
        """Finished contents (no remaining use) swap into their home slot when that slot is in
        the tile and its current assignee is not parked there for the next pass."""
        by_slot = {p: c for c, p in asg.items()}
        for c in content:
            if c not in finished or asg[c] == home[c] or home[c] not in by_slot:
                continue
            other = by_slot[home[c]]
            if other in parked or (other in finished and asg[other] == home[other]):
                continue
            if home[other] >= self.n_local and asg[other] >= self.n_local - (self.n - self.n_local):
                continue          # `other` waits on a top slot for the swap that takes it home
            if c in parked:
                continue
            asg[c], asg[other] = asg[other], asg[c]
            by_slot[asg[c]], by_slot[asg[other]] = c, other
        # a finished content whose home is a RANK bit leaves with the next swap: lift it out of the
        # low positions (a swap names local positions >= 5, else it needs a relabel pass first)
        if self.n_local < self.n:
            for c in content:
                if c in finished and home[c] >= self.n_local and asg[c] < 5 and c not in parked:
                    cand = [d for d in content if asg[d] >= 5 and d not in parked
                            and not (home[d] >= self.n_local and d in finished)
                            and not (d in finished and asg[d] == home[d])]
                    if cand:
                        d = max(cand, key=lambda x: asg[x])
                        asg[c], asg[d] = asg[d], asg[c]
        return asg

    def _encode_tphase(self, op: MicroOp, slot_of, idx_of, pos, thr, tables: list) -> L.QsvOp:
        """OP_TPHASE record + its tables: one per-thread table for the partners that are thread-fixed
        tile bits, one 256-entry table per 8-bit run of the global index for the partners outside
        the tile (rank bits included)."""
        t = self.t
        o = L.QsvOp()
        o.kind = L.OP_TPHASE
        o.target = slot_of[op.target]
        off = sum(len(x) for x in tables)
        thr_part = {idx_of[c]: ph for c, ph in op.tph.items() if c in idx_of}
        glob_part = {pos[c]: ph for c, ph in op.tph.items() if c not in idx_of}
        assert not any(c in slot_of for c in op.tph)
        o.m[0], o.m[1], o.m[2], o.m[3] = -1.0, -1.0, 0.0, 0.0
        if thr_part:
            tix = np.arange(1 << (t - REG_BITS))
            tab = np.ones(len(tix), dtype=np.complex128)
            for k, i in enumerate(thr):
                if i in thr_part:
                    tab[((tix >> k) & 1) == 1] *= thr_part[i]
            tables.append(tab / np.abs(tab))
            o.m[0] = float(off)
            off += len(tab)
        if glob_part:
            runs = sorted({p_ >> 3 for p_ in glob_part})
            if runs[-1] > 7:
                raise NotImplementedError("table phase on an index bit >= 64")
            b = np.arange(256)
            o.m[1] = float(off)
            mask = 0
            for r in runs:
                tab = np.ones(256, dtype=np.complex128)
                for p_, ph in glob_part.items():
                    if p_ >> 3 == r:
                        tab[((b >> (p_ & 7)) & 1) == 1] *= ph
                tables.append(tab / np.abs(tab))
                mask |= 1 << r
            o.m[2] = float(mask)
        return o

    def _encode(self, op: MicroOp, slot_of, idx_of, pos) -> L.QsvOp:
        o = L.QsvOp()
        o.kind = op.kind
        if op.target is not None:
            o.target = slot_of[op.target]
        reg_ctrl = tile_ctrl = glob = 0
        for c in op.ctrls:
            if c in slot_of:
                reg_ctrl |= 1 << slot_of[c]
            elif c in idx_of:
                tile_ctrl |= 1 << idx_of[c]
            else:
                glob |= 1 << pos[c]
        o.reg_ctrl, o.tile_ctrl, o.glob_ctrl = reg_ctrl, tile_ctrl, glob
        for k in range(4):
            o.m[k] = op.m[k]
        if op.kind in (L.OP_HAD, L.OP_ROT):
            o.m[2] = o.m[3] = 0.0
            flags, neg = 0, bool(op.pre_neg)
            ph = complex(op.pre_phase)
            if ph != _ONE:
                ph /= abs(ph)
                if abs(ph.imag) <= _SNAP:
                    ph = complex(1.0 if ph.real > 0 else -1.0, 0.0)
                if ph.real < 0:                     # keep |phi| <= pi/2: e^{i phi} = -e^{i (phi -+ pi)}
                    ph, neg = -ph, not neg
                if ph != _ONE:
                    flags |= L.OPF_PREPHASE
                    o.m[2], o.m[3] = ph.imag / (1.0 + ph.real), ph.imag
            if neg:
                flags |= L.OPF_PRENEG
            if op.pre_par:
                assert not op.ctrls
                flags |= L.OPF_PRESIGN
                for c in op.pre_par:
                    assert c not in slot_of
                    if c in idx_of:
                        o.tile_ctrl |= 1 << idx_of[c]
                    else:
                        o.glob_ctrl |= 1 << pos[c]
            o.flags = flags
        return o

    # ---- global <-> local swaps ---------------------------------------------------------
    def _relabel_to(self, prog, pos, home, xf, want: dict, touch=()) -> None:
        """Relabel-only passes that move content c to local position want[c] (the displaced
        contents take the vacated slots).  Flips of finished contents are materialised; contents
        in `touch` are put through a pass even if they are already in place."""
        todo = {c: p for c, p in want.items() if pos[c] != p or c in touch}
        while todo:
            at = {p: c for c, p in enumerate(pos)}
            chosen = set(range(self.a))
            batch = {}
            for c, p in todo.items():
                extra = {pos[c], p} - chosen
                if len(chosen) + len(extra) > self.t:
                    continue
                chosen |= extra
                batch[c] = p
            if not batch:
                raise RuntimeError("relabel: tile too small")
            for p in range(self.n_local):
                if len(chosen) >= self.t:
                    break
                chosen.add(p)
            tile = [at[p] for p in sorted(chosen)]
            # explicit store: movers go to their slots; the evicted take the movers' old slots
            new = {c: pos[c] for c in tile}
            for c, p in batch.items():
                d = next(x for x in tile if new[x] == p)
                new[d], new[c] = new[c], p
            prog.steps.append(self._emit_pass(tile, [], pos, home, [], False, xf, explicit_store=new))
            todo = {c: p for c, p in todo.items() if c not in batch}

    def _do_swap(self, prog, pos, home, xf, incoming: list, outgoing: list, materialise=False) -> None:
        """incoming[i] (on a rank bit) <-> outgoing[i] (local): park the outgoing contents on the
        top local positions, then one all-to-all."""
        s_ = len(incoming)
        assert s_ == len(outgoing) and s_ > 0
        top = [self.n_local - s_ + i for i in range(s_)]
        touch = {c for c in outgoing if xf[c] and self._uses[c] == 0} if materialise else set()
        if self.swap_anywhere and not touch and all(pos[c] >= 5 for c in outgoing):
            gbits, lbits = [pos[c] for c in incoming], [pos[c] for c in outgoing]
            prog.steps.append(SwapStep(gbits, lbits))
            for i in range(s_):
                pos[incoming[i]], pos[outgoing[i]] = lbits[i], gbits[i]
            return
        # any pairing of rank bits with top slots is one all-to-all: keep outgoing contents that
        # already sit on a top slot where they are (saves the relabel pass)
        pairs = list(zip(incoming, outgoing))
        order = [None] * s_
        for pr in list(pairs):
            if pos[pr[1]] in top and order[top.index(pos[pr[1]])] is None:
                order[top.index(pos[pr[1]])] = pr
                pairs.remove(pr)
        for j in range(s_):
            if order[j] is None:
                order[j] = pairs.pop(0)
        incoming, outgoing = [p_[0] for p_ in order], [p_[1] for p_ in order]
        self._relabel_to(prog, pos, home, xf, {c: top[i] for i, c in enumerate(outgoing)}, touch)
        gbits = [pos[c] for c in incoming]
        prog.steps.append(SwapStep(gbits, top))
        for i in range(s_):
            pos[incoming[i]], pos[outgoing[i]] = top[i], gbits[i]

    def _swap_in(self, remaining, pos, home, prog, xf) -> None:
        """The stage is exhausted: bring every rank-bit content that still has to be MIXED into the
        shard, sending out the local contents that need it least (finished first, then those
        whose next use as a target is furthest away; a content whose home is a rank bit is
        preferred so that the final restoration may need no swap)."""
        first_target: dict = {}
        for i, op in enumerate(remaining):
            if op.target is not None and op.target not in first_target:
                first_target[op.target] = i
        incoming = sorted((c for c in first_target if pos[c] >= self.n_local), key=lambda c: first_target[c])
        if not incoming:
            raise RuntimeError("swap requested but no rank-bit content needs mixing")
        inf = len(remaining) + 1

        def key(c):
            return (-first_target.get(c, inf), 0 if home[c] >= self.n_local else 1, -pos[c])

        locals_ = sorted((c for c in range(self.n) if pos[c] < self.n_local), key=key)
        chosen = locals_[: len(incoming)]
        # send a content whose home is one of the vacated rank bits exactly there
        outgoing = [None] * len(incoming)
        for i, cin in enumerate(incoming):
            hit = next((c for c in chosen if home[c] == pos[cin]), None)
            if hit is not None:
                outgoing[i] = hit
                chosen.remove(hit)
        for i in range(len(incoming)):
            if outgoing[i] is None:
                outgoing[i] = chosen.pop(0)
        # a flip pending on an UNFINISHED outgoing content simply stays in the frame (xf); a finished
        # one is materialised by the relabel pass now (it could not be undone on a rank bit)
        self._do_swap(prog, pos, home, xf, incoming, outgoing,
                      materialise=not self.rank_flips and any(xf[c] and self._uses[c] == 0 for c in outgoing))

    def _restore_global(self, prog, pos, home, xf) -> None:
        """After the last op: contents whose home is a rank bit go back out, the exiles on rank
        bits come home.  A pending X on a rank bit is brought into the shard first (it can
        only be materialised as a store-address flip)."""
        n_loc = self.n_local
        flipped_glob = [c for c in range(self.n) if pos[c] >= n_loc and xf[c] and home[c] == pos[c]]
        if flipped_glob and not self.rank_flips:
            locals_ = sorted((c for c in range(self.n) if pos[c] < n_loc), key=lambda c: -pos[c])
            self._do_swap(prog, pos, home, xf, flipped_glob, locals_[: len(flipped_glob)])
        for _ in range(4):
            out = [c for c in range(self.n) if pos[c] < n_loc and home[c] >= n_loc]
            inc = [c for c in range(self.n) if pos[c] >= n_loc and home[c] != pos[c]]
            if not out and not inc:
                return
            # every rank bit that is someone's unmet home, or holds a displaced content, takes part;
            # the content whose home it is goes there if it is local, else a filler that comes
            # back in the next iteration (a rank-bit -> rank-bit move needs two swaps)
            at = {p: c for c, p in enumerate(pos)}
            bits = sorted({home[c] for c in out} | {pos[c] for c in inc})
            incoming = [at[b] for b in bits]
            out_by_home = {home[c]: c for c in out}
            fillers = [c for c in sorted(range(self.n), key=lambda c: -pos[c])
                       if pos[c] < n_loc and home[c] < n_loc]
            outgoing = [out_by_home[b] if b in out_by_home else fillers.pop(0) for b in bits]
            # all ops are done (uses == 0): the relabel pass materialises the flips of what leaves
            self._do_swap(prog, pos, home, xf, incoming, outgoing, materialise=not self.rank_flips)
        raise RuntimeError("rank-bit restoration did not converge")

    # ---- layout restoration -----------------------------------------------------------
    def _restore(self, prog, pos, home, xf) -> None:
        """Relabel-only passes until every local content is at its home position."""
        n_loc = self.n_local
        if n_loc < self.n and self.allow_swaps:
            self._restore_global(prog, pos, home, xf)
        for _ in range(8 * self.n + 8):
            bad = [c for c in range(self.n) if pos[c] < n_loc and pos[c] != home[c]]
            flipped = [c for c in range(self.n) if xf[c]]
            if self.rank_flips:
                flipped = [c for c in flipped if pos[c] < n_loc]      # rank-bit flips rename the shards
            if any(pos[c] >= n_loc for c in flipped):
                raise NotImplementedError("a pending X sits on a rank bit")
            if not bad and not flipped:
                return
            if any(home[c] >= n_loc for c in bad):
                raise NotImplementedError("layout restoration across rank bits needs a global swap")
            at = {p: c for c, p in enumerate(pos)}
            # cycles of the position permutation p -> home[at[p]]
            cycles, seen = [], set()
            for p in sorted(pos[c] for c in bad):
                if p in seen:
                    continue
                cyc, x = [], p
                while x not in seen:
                    seen.add(x)
                    cyc.append(x)
                    x = home[at[x]]
                cycles.append(cyc)
            cycles.sort(key=lambda cyc: (not any(p < self.a for p in cyc), len(cyc)))
            chosen = set(range(self.a))
            for c in flipped:                              # pending flips need their bit in a tile
                if len(chosen) < self.t:
                    chosen.add(pos[c])
            for cyc in cycles:
                fresh = [p for p in cyc if p not in chosen]
                room = self.t - len(chosen)
                if len(fresh) <= room:
                    chosen.update(fresh)
                elif room >= 2:
                    # rotate so the chain starts right after a position already chosen, if any
                    k = next((i for i, p in enumerate(cyc) if p in chosen), -1)
                    chain = cyc[k + 1:] + cyc[:k + 1]
                    chosen.update([p for p in chain if p not in chosen][:room])
                if len(chosen) >= self.t:
                    break
            for p in range(n_loc):                         # pad with idle positions (lowest first)
                if len(chosen) >= self.t:
                    break
                chosen.add(p)
            tile = [at[p] for p in sorted(chosen)]
            prog.steps.append(self._emit_pass(tile, [], pos, home, [], True, xf))
        raise RuntimeError("layout restoration did not converge")


def compile_ops(ir_ops, n_qubits: int, n_local: int | None = None, dtype: str = "complex128",
                **kw) -> Program:
    return PassCompiler(n_qubits, n_local, dtype, **kw).compile(ir_ops)
