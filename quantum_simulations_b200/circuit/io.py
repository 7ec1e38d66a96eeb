"""Circuit-dict contract: validation, normalisation and ASAP levelization.

Host-side mirror of the reference contract (wenbo_engine/circuit/io.py:12-117,
wenbo_engine/docs/circuit_contract.md).  Same public names, same normalised output
(``{"number_of_qubits": n, "gates": [{"qubits", "gate", "params"}]}``) and the same
``ValueError`` message stems the reference tests match on
(tests/test_contract.py:28-78): "missing required keys", "unknown top-level",
"unsupported gate", "needs N qubit(s)", "out of range", "requires param",
"unknown keys".

Index convention (reference io.py:3-6): LITTLE-ENDIAN — qubit q is bit q of the
amplitude index.
"""
from __future__ import annotations

import re
from typing import Any

ENDIANNESS = "little"

# name -> (arity, {param: python type | "array"})
_GATE_TABLE: dict[str, tuple[int, dict[str, Any]]] = {
    "H": (1, {}), "X": (1, {}), "Y": (1, {}), "Z": (1, {}), "S": (1, {}), "T": (1, {}),
    "RY": (1, {"theta": float}),
    "R": (1, {"k": int}),
    "G": (1, {"p": int}),
    "CNOT": (2, {}), "SWAP": (2, {}), "CZ": (2, {}), "CY": (2, {}),
    "CR": (2, {"k": int}),
    "CU": (2, {"U": "array", "exponent": int}),
}

GATES_1Q_NO_PARAMS = frozenset(g for g, (a, p) in _GATE_TABLE.items() if a == 1 and not p)
GATES_2Q_NO_PARAMS = frozenset(g for g, (a, p) in _GATE_TABLE.items() if a == 2 and not p)
GATES_1Q_PARAM_SPEC = {g: p for g, (a, p) in _GATE_TABLE.items() if a == 1 and p}
GATES_2Q_PARAM_SPEC = {g: p for g, (a, p) in _GATE_TABLE.items() if a == 2 and p}
ALL_1Q = frozenset(g for g, (a, _) in _GATE_TABLE.items() if a == 1)
ALL_2Q = frozenset(g for g, (a, _) in _GATE_TABLE.items() if a == 2)
ALL_GATES = ALL_1Q | ALL_2Q

_TOP_KEYS = {"number_of_qubits", "gates"}
_GATE_KEYS = {"qubits", "gate", "params"}
_NAME_ENCODED = re.compile(r"^(CR|R)(\d+)$")


def _parse_name_encoded(raw: str) -> tuple[str, dict]:
    """'CR3' -> ('CR', {'k': 3}); 'R5' -> ('R', {'k': 5}); anything else unchanged.

    Reference behaviour: io.py:32-41 ('RY' is never read as R with a suffix since the
    suffix must be digits)."""
    hit = _NAME_ENCODED.match(raw) if isinstance(raw, str) else None
    if hit:
        return hit.group(1), {"k": int(hit.group(2))}
    return raw, {}


def _normalise_gate(entry: Any, n_qubits: int, pos: int) -> dict:
    where = f"gate[{pos}]"
    if not isinstance(entry, dict):
        raise ValueError(f"{where}: must be a dict")
    keys = set(entry)
    if "qubits" not in keys or "gate" not in keys:
        raise ValueError(f"{where}: missing 'qubits' or 'gate'")
    stray = keys - _GATE_KEYS
    if stray:
        raise ValueError(f"{where}: unknown keys {stray}")

    qubits = entry["qubits"]
    if not isinstance(qubits, list) or any(not isinstance(q, int) for q in qubits):
        raise ValueError(f"{where}: qubits must be list[int]")
    for q in qubits:
        if not 0 <= q < n_qubits:
            raise ValueError(f"{where}: qubit {q} out of range [0, {n_qubits})")

    name, implied = _parse_name_encoded(entry["gate"])
    if name not in _GATE_TABLE:
        raise ValueError(f"{where}: unsupported gate '{entry['gate']}'")
    arity, spec = _GATE_TABLE[name]
    if len(qubits) != arity:
        raise ValueError(f"{where}: {name} needs {arity} qubit(s), got {len(qubits)}")

    params = dict(implied)
    params.update(entry.get("params") or {})  # explicit params win over name-encoded
    for pname, ptype in spec.items():
        if pname not in params:
            raise ValueError(f"{where}: {name} requires param '{pname}'")
        if ptype != "array" and not isinstance(params[pname], (ptype, int)):
            raise ValueError(f"{where}: param '{pname}' bad type")
    return {"qubits": list(qubits), "gate": name, "params": params}


def validate_circuit_dict(d: dict[str, Any]) -> dict:
    """Validate and normalise a circuit dict; raises ValueError on bad input."""
    if not isinstance(d, dict):
        raise ValueError("circuit must be a dict")
    missing = _TOP_KEYS - set(d)
    if missing:
        raise ValueError(f"missing required keys: {missing}")
    extra = set(d) - _TOP_KEYS
    if extra:
        raise ValueError(f"unknown top-level keys: {extra}")
    n = d["number_of_qubits"]
    if not isinstance(n, int) or n < 1:
        raise ValueError(f"number_of_qubits must be positive int, got {n!r}")
    gates = d["gates"]
    if not isinstance(gates, list):
        raise ValueError("gates must be a list")
    return {
        "number_of_qubits": n,
        "gates": [_normalise_gate(g, n, i) for i, g in enumerate(gates)],
    }


def levelize(circuit_dict: dict) -> list[list[dict]]:
    """ASAP levels: a gate goes to the first level at which all its qubits are free.

    This defines the "gate layer" of the throughput metric (reference io.py:106-117).
    """
    next_free: dict[int, int] = {}
    levels: list[list[dict]] = []
    for gate in circuit_dict["gates"]:
        lvl = 0
        for q in gate["qubits"]:
            lvl = max(lvl, next_free.get(q, 0))
        if lvl >= len(levels):
            levels.extend([] for _ in range(lvl + 1 - len(levels)))
        levels[lvl].append(gate)
        for q in gate["qubits"]:
            next_free[q] = lvl + 1
    return levels
