"""OpenQASM 2.0 front end for the circuits vendored with the reference
(v3_hisvsim_spark/hisvsim_repo/QASMBench/cluster/*/*.qasm: qft_n28, cat_state_n30, bv_n30, ...;
SURVEY.md section 8f-3).  Output is the step IR the pass compiler consumes: ``(n_qubits, [(qubits, U)])``
with 2x2 / 4x4 complex128 matrices, qubits little-endian like the circuit dict
(reference circuit/io.py:3-6).  `qasm_to_dict` additionally maps a program onto the reference's
named gate set when every gate has a name there.

Supported: qreg (several, concatenated in declaration order), the qelib1 gates
    id x y z h s sdg t tdg sx rx ry rz p u1 u2 u3 u   cx cy cz ch swap cp cu1 crz cu3 rzz rxx ryy   ccx cswap
user `gate` definitions (expanded), broadcast over whole registers, `barrier` / `creg` / `measure`
(ignored: the engine returns the state).  `reset` and `if` are rejected (not unitary / classical)."""
from __future__ import annotations

import ast
import math
import operator
import re

import numpy as np

_S2 = 1.0 / math.sqrt(2.0)
C128 = np.complex128


def _u3(theta, phi, lam):
    c, s = math.cos(theta / 2), math.sin(theta / 2)
    return np.array([[c, -np.exp(1j * lam) * s], [np.exp(1j * phi) * s, np.exp(1j * (phi + lam)) * c]], dtype=C128)


_FIXED_1Q = {
    "id": np.eye(2, dtype=C128), "x": np.array([[0, 1], [1, 0]], dtype=C128),
    "y": np.array([[0, -1j], [1j, 0]], dtype=C128), "z": np.diag([1, -1]).astype(C128),
    "h": np.array([[1, 1], [1, -1]], dtype=C128) * _S2, "s": np.diag([1, 1j]).astype(C128),
    "sdg": np.diag([1, -1j]).astype(C128), "t": np.diag([1, np.exp(0.25j * math.pi)]).astype(C128),
    "tdg": np.diag([1, np.exp(-0.25j * math.pi)]).astype(C128),
    "sx": 0.5 * np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]], dtype=C128),
}
_PARAM_1Q = {
    "rx": lambda t: np.array([[math.cos(t / 2), -1j * math.sin(t / 2)], [-1j * math.sin(t / 2), math.cos(t / 2)]], dtype=C128),
    "ry": lambda t: np.array([[math.cos(t / 2), -math.sin(t / 2)], [math.sin(t / 2), math.cos(t / 2)]], dtype=C128),
    "rz": lambda t: np.diag([np.exp(-0.5j * t), np.exp(0.5j * t)]).astype(C128),
    "p": lambda l: np.diag([1, np.exp(1j * l)]).astype(C128),
    "u1": lambda l: np.diag([1, np.exp(1j * l)]).astype(C128),
    "u2": lambda p, l: _u3(math.pi / 2, p, l),
    "u3": _u3, "u": _u3,
}
# dict-gate names of the reference (kernel/gates.py) for qasm_to_dict
_DICT_NAMES = {"x": "X", "y": "Y", "z": "Z", "h": "H", "s": "S", "t": "T", "cx": "CNOT", "cz": "CZ", "cy": "CY",
               "swap": "SWAP"}


def _controlled(u):
    m = np.eye(4, dtype=C128)                    # row = 2*bit(control) + bit(target): control = qubits[0]
    m[2:, 2:] = u
    return m


_SWAP = np.eye(4, dtype=C128)[[0, 2, 1, 3]]


class QasmError(ValueError):
    pass


_FUNCS = {"sin": math.sin, "cos": math.cos, "tan": math.tan, "exp": math.exp, "ln": math.log, "sqrt": math.sqrt}
_BINOPS = {ast.Add: operator.add, ast.Sub: operator.sub, ast.Mult: operator.mul, ast.Div: operator.truediv,
           ast.Pow: operator.pow}


def _eval(expr: str, env: dict) -> float:
    """Parameter expression of OpenQASM 2.0 (numbers, pi, formal parameters, + - * / ^, unary minus, and the
    six built-in functions).  The text is parsed to a Python AST and walked over a closed set of node types:
    attribute access, subscripts, calls of anything but the six functions, comparisons etc. are rejected, so a
    .qasm file cannot reach interpreter internals (nothing is ever passed to eval())."""
    expr = expr.strip()
    try:
        tree = ast.parse(expr.replace("^", "**"), mode="eval")
    except (SyntaxError, ValueError, MemoryError, RecursionError):
        raise QasmError(f"bad parameter expression {expr!r}") from None

    def walk(node, depth=0):
        if depth > 64:
            raise QasmError(f"parameter expression {expr!r} is nested too deeply")
        if isinstance(node, ast.Expression):
            return walk(node.body, depth + 1)
        if isinstance(node, ast.Constant) and type(node.value) in (int, float):
            return float(node.value)
        if isinstance(node, ast.Name):
            if node.id == "pi":
                return math.pi
            if node.id in env:
                return float(env[node.id])
            raise QasmError(f"unknown name {node.id!r} in parameter expression {expr!r}")
        if isinstance(node, ast.UnaryOp) and isinstance(node.op, (ast.USub, ast.UAdd)):
            v = walk(node.operand, depth + 1)
            return -v if isinstance(node.op, ast.USub) else v
        if isinstance(node, ast.BinOp) and type(node.op) in _BINOPS:
            a, b = walk(node.left, depth + 1), walk(node.right, depth + 1)
            if isinstance(node.op, ast.Pow) and abs(b) > 1024:
                raise QasmError(f"exponent too large in {expr!r}")
            return float(_BINOPS[type(node.op)](a, b))
        if (isinstance(node, ast.Call) and isinstance(node.func, ast.Name) and node.func.id in _FUNCS
                and len(node.args) == 1 and not node.keywords):
            return float(_FUNCS[node.func.id](walk(node.args[0], depth + 1)))
        raise QasmError(f"bad parameter expression {expr!r}: {type(node).__name__} is not allowed")

    try:
        return walk(tree)
    except QasmError:
        raise
    except (ArithmeticError, ValueError, TypeError) as e:
        raise QasmError(f"cannot evaluate {expr!r}: {e}") from None


def _split_args(s: str) -> list[str]:
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            depth += ch == "("
            depth -= ch == ")"
            cur += ch
    if cur.strip():
        out.append(cur)
    return [a.strip() for a in out]


def qasm_to_ops(text: str, with_statement_index: bool = False, _events: list | None = None, _cregs: dict | None = None):
    """Parse an OpenQASM 2.0 program -> (n_qubits, [(qubits, U)]) in program order.
    with_statement_index=True adds a third value: for every op the 0-based index of the gate
    STATEMENT it came from (a broadcast or a ccx yields several ops of one statement; barrier /
    measure / declarations are not counted) — what HiSVSIM's part files number.
    (_events / _cregs: used by qasm_to_steps, which keeps measure / reset / if instead of refusing them.)"""
    text = re.sub(r"//[^\n]*", "", text)
    regs: dict[str, tuple[int, int]] = {}           # name -> (offset, size)
    gates: dict[str, tuple[list, list, list]] = {}  # name -> (params, qargs, body statements)
    ops: list = []
    stmt_of: list = []
    n_stmt = 0
    n = 0
    touched: set = set()                            # qubits some gate has acted on (for leading resets)
    scanned = [0]

    # pull out gate definitions first (they contain braces)
    def take_gate(m):
        head, body = m.group(1), m.group(2)
        hm = re.fullmatch(r"\s*(\w+)\s*(?:\(([^)]*)\))?\s*([\w\s,]*)", head)
        if not hm:
            raise QasmError(f"bad gate definition {head!r}")
        params = [p.strip() for p in (hm.group(2) or "").split(",") if p.strip()]
        qargs = [q.strip() for q in hm.group(3).split(",") if q.strip()]
        gates[hm.group(1)] = (params, qargs, [s.strip() for s in body.split(";") if s.strip()])
        return ""

    text = re.sub(r"\bgate\b([^{]*)\{([^}]*)\}", take_gate, text)

    def emit(name: str, params: list[float], qs: list[int]) -> None:
        if len(set(qs)) != len(qs):
            raise QasmError(f"{name}: repeated qubit {qs}")
        if name in _FIXED_1Q:
            ops.append(([qs[0]], _FIXED_1Q[name]))
        elif name in _PARAM_1Q:
            ops.append(([qs[0]], _PARAM_1Q[name](*params)))
        elif name in ("cx", "CX"):
            ops.append(([qs[0], qs[1]], _controlled(_FIXED_1Q["x"])))
        elif name in ("cy", "cz", "ch"):
            ops.append(([qs[0], qs[1]], _controlled(_FIXED_1Q[name[1]])))
        elif name in ("cp", "cu1"):
            ops.append(([qs[0], qs[1]], _controlled(_PARAM_1Q["u1"](*params))))
        elif name == "crz":
            ops.append(([qs[0], qs[1]], _controlled(_PARAM_1Q["rz"](*params))))
        elif name == "cu3":
            ops.append(([qs[0], qs[1]], _controlled(_u3(*params))))
        elif name == "swap":
            ops.append(([qs[0], qs[1]], _SWAP))
        elif name == "rzz":                          # exp(-i t/2 Z(x)Z): diagonal
            t = params[0]
            ops.append(([qs[0], qs[1]], np.diag([np.exp(-0.5j * t), np.exp(0.5j * t), np.exp(0.5j * t), np.exp(-0.5j * t)]).astype(C128)))
        elif name in ("rxx", "ryy"):                 # exp(-i t/2 P(x)P), P = X or Y, as (W(x)W) RZZ(t) (W(x)W)^-1
            # with W Z W^-1 = +-P: 1-qubit gates around a DIAGONAL 2-qubit gate, which the pass compiler
            # runs inside fused passes (a dense non-controlled 4x4 would cost a sweep of its own)
            before, after = (("h", []), ("h", [])) if name == "rxx" else (("rx", [-math.pi / 2]), ("rx", [math.pi / 2]))
            for q in qs:
                emit(before[0], before[1], [q])
            emit("rzz", params, qs)
            for q in qs:
                emit(after[0], after[1], [q])
        elif name == "ccx":                          # standard 6-CNOT decomposition (qelib1.inc)
            a, b, c = qs
            for g_, q_ in (("h", [c]), ("cx", [b, c]), ("tdg", [c]), ("cx", [a, c]), ("t", [c]), ("cx", [b, c]),
                           ("tdg", [c]), ("cx", [a, c]), ("t", [b]), ("t", [c]), ("h", [c]), ("cx", [a, b]),
                           ("t", [a]), ("tdg", [b]), ("cx", [a, b])):
                emit(g_, [], q_)
        elif name == "cswap":
            a, b, c = qs
            emit("cx", [], [c, b]); emit("ccx", [], [a, b, c]); emit("cx", [], [c, b])
        elif name in gates:
            gp, gq, body = gates[name]
            if len(gp) != len(params) or len(gq) != len(qs):
                raise QasmError(f"{name}: expected {len(gp)} parameters and {len(gq)} qubits")
            penv, qenv = dict(zip(gp, params)), dict(zip(gq, qs))
            for st in body:
                bm = re.fullmatch(r"(\w+)\s*(?:\((.*)\))?\s*(.*)", st)
                if not bm or bm.group(1) == "barrier":
                    continue
                bparams = [_eval(e, penv) for e in _split_args(bm.group(2) or "")]
                emit(bm.group(1), bparams, [qenv[a.strip()] for a in bm.group(3).split(",") if a.strip()])
        else:
            raise QasmError(f"unsupported gate '{name}'")

    for st in [s.strip() for s in text.split(";")]:
        if not st or st.startswith("OPENQASM") or st.startswith("include"):
            continue
        m = re.fullmatch(r"qreg\s+(\w+)\s*\[\s*(\d+)\s*\]", st)
        if m:
            regs[m.group(1)] = (n, int(m.group(2)))
            n += int(m.group(2))
            continue
        if _events is not None:                    # classical mode: keep the non-unitary statements as events
            m = re.fullmatch(r"creg\s+(\w+)\s*\[\s*(\d+)\s*\]", st)
            if m:
                _cregs[m.group(1)] = int(m.group(2))
                continue
            m = re.fullmatch(r"measure\s+(\w+)\s*(?:\[\s*(\d+)\s*\])?\s*->\s*(\w+)\s*(?:\[\s*(\d+)\s*\])?", st)
            if m:
                if m.group(1) not in regs or m.group(3) not in _cregs:
                    raise QasmError(f"measure: unknown register in {st!r}")
                off, size = regs[m.group(1)]
                qs_ = range(off, off + size) if m.group(2) is None else [off + int(m.group(2))]
                bs_ = range(_cregs[m.group(3)]) if m.group(4) is None else [int(m.group(4))]
                if len(qs_) != len(bs_) or any(b >= _cregs[m.group(3)] for b in bs_):
                    raise QasmError(f"measure: register sizes do not match in {st!r}")
                for q_, b_ in zip(qs_, bs_):
                    _events.append((len(ops), "measure", q_, m.group(3), b_))
                continue
            m = re.fullmatch(r"reset\s+(\w+)\s*(?:\[\s*(\d+)\s*\])?", st)
            if m and m.group(1) in regs:
                off, size = regs[m.group(1)]
                for q_ in (range(off, off + size) if m.group(2) is None else [off + int(m.group(2))]):
                    _events.append((len(ops), "reset", q_))
                continue
            m = re.fullmatch(r"if\s*\(\s*(\w+)\s*==\s*(\d+)\s*\)\s*(.*)", st, re.S)
            if m:
                if m.group(1) not in _cregs:
                    raise QasmError(f"if: unknown classical register {m.group(1)!r}")
                if re.match(r"(measure|reset|if)\b", m.group(3)):
                    raise QasmError("if: only gate statements can be conditional")
                before_if = len(ops)
                conditional = ("if", m.group(1), int(m.group(2)))
                st = m.group(3).strip()                     # falls through: the gate statement is parsed below
            else:
                conditional = None
        else:
            conditional = None
        if re.match(r"(creg|barrier|measure)\b", st):
            continue
        m = re.fullmatch(r"reset\s+(\w+)\s*(?:\[\s*(\d+)\s*\])?", st)
        if m and m.group(1) in regs:
            # reset of a qubit no gate has acted on yet: the run starts from |0...0>, so it is the identity
            # (QASMBench's bwt / square_root open with such resets); anywhere else it is not unitary
            off, size = regs[m.group(1)]
            which = range(off, off + size) if m.group(2) is None else [off + int(m.group(2))]
            for qs_, _ in ops[scanned[0]:]:                 # incremental: every op is looked at once
                touched.update(qs_)
            scanned[0] = len(ops)
            if any(q in touched for q in which):
                raise QasmError("'reset' after a gate on the same qubit is not supported (not a unitary gate)")
            continue
        if re.match(r"(reset|if)\b", st):
            raise QasmError(f"'{st.split()[0]}' is not supported (not a unitary gate)")
        m = re.fullmatch(r"(\w+)\s*(?:\((.*)\))?\s+(.*)", st, re.S)
        if not m:
            raise QasmError(f"cannot parse statement {st!r}")
        params = [_eval(e, {}) for e in _split_args(m.group(2) or "")]
        args = []
        for a in m.group(3).split(","):
            am = re.fullmatch(r"\s*(\w+)\s*(?:\[\s*(\d+)\s*\])?\s*", a)
            if not am or am.group(1) not in regs:
                raise QasmError(f"unknown qubit argument {a.strip()!r}")
            off, size = regs[am.group(1)]
            if am.group(2) is None:
                args.append([off + i for i in range(size)])           # whole register: broadcast
            else:
                if int(am.group(2)) >= size:
                    raise QasmError(f"{a.strip()}: index out of range")
                args.append([off + int(am.group(2))])
        width = max(len(a) for a in args)
        before = len(ops)
        for i in range(width):
            emit(m.group(1), params, [a[i] if len(a) > 1 else a[0] for a in args])
        stmt_of += [n_stmt] * (len(ops) - before)
        n_stmt += 1
        if conditional is not None:                 # the ops of this statement leave the unconditional stream
            cond_ops = ops[before_if:]
            del ops[before_if:]
            del stmt_of[before_if:]
            _events.append((len(ops), conditional[0], conditional[1], conditional[2], cond_ops))
    if n == 0:
        raise QasmError("no qreg declared")
    return (n, ops, stmt_of) if with_statement_index else (n, ops)


def qasm_to_steps(text: str):
    """OpenQASM 2.0 WITH its non-unitary statements: mid-circuit `measure`, `reset`, and classically controlled
    gates `if(c==k) gate ...;` (six of the QASMBench circuits vendored with the reference use them).
    Returns (n_qubits, steps, cregs): cregs = {name: size}; steps is a list of
        ("ops", [(qubits, U), ...])            a unitary segment (as qasm_to_ops)
        ("measure", qubit, creg, bit)          projective Z measurement, outcome -> creg[bit]
        ("reset", qubit)                       measure, then X if the outcome was 1
        ("if", creg, value, [(qubits, U)...])  the ops run iff the integer value of creg (bit 0 = LSB) == value
    Execution semantics (one TRAJECTORY under a seed) are frozen in oracle/ref_dense.py::run_qasm_steps; the
    reference has no such path (its converter drops measurements), HiSVSIM's collapse primitive is
    state_vector.hpp:829-893."""
    events: list = []
    cregs: dict = {}
    n, ops = qasm_to_ops(text, _events=events, _cregs=cregs)
    steps: list = []
    at = 0
    for ev in events:
        pos = ev[0]
        if pos > at:
            steps.append(("ops", ops[at:pos]))
            at = pos
        steps.append(tuple(ev[1:]))
    if at < len(ops):
        steps.append(("ops", ops[at:]))
    return n, steps, cregs


def is_unitary_program(steps) -> bool:
    """True if the steps can run as ONE unitary circuit with the measurements read at the end: no reset, no if,
    and no gate on a qubit after that qubit was measured."""
    measured: set = set()
    for st in steps:
        if st[0] in ("reset", "if"):
            return False
        if st[0] == "measure":
            measured.add(st[1])
        elif st[0] == "ops" and any(q in measured for qs, _ in st[1] for q in qs):
            return False
    return True


def qasm_to_dict(text: str) -> dict:
    """Circuit dict in the reference's format when every gate has a NAME there (H X Y Z S T RY CNOT CZ
    CY SWAP, reference kernel/gates.py:24-108); other programs need `qasm_to_ops`."""
    gates = []
    text_nc = re.sub(r"//[^\n]*", "", text)
    n = sum(int(s) for s in re.findall(r"qreg\s+\w+\s*\[\s*(\d+)\s*\]", text_nc))
    regs, off = {}, 0
    for name, size in re.findall(r"qreg\s+(\w+)\s*\[\s*(\d+)\s*\]", text_nc):
        regs[name] = off
        off += int(size)
    for st in [s.strip() for s in text_nc.split(";")]:
        m = re.fullmatch(r"(\w+)\s*(?:\((.*)\))?\s+((?:\w+\s*\[\s*\d+\s*\]\s*,?\s*)+)", st)
        if not m or m.group(1) in ("qreg", "creg", "barrier", "measure"):
            continue
        qs = [regs[r] + int(i) for r, i in re.findall(r"(\w+)\s*\[\s*(\d+)\s*\]", m.group(3))]
        name = m.group(1)
        if name == "ry":
            gates.append({"qubits": qs, "gate": "RY", "params": {"theta": _eval(m.group(2), {})}})
        elif name in _DICT_NAMES:
            gates.append({"qubits": qs, "gate": _DICT_NAMES[name], "params": {}})
        else:
            raise QasmError(f"gate '{name}' has no name in the reference's circuit dict; use qasm_to_ops")
    return {"number_of_qubits": n, "gates": gates}
