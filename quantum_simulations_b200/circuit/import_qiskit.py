"""Qiskit QuantumCircuit -> circuit dict (reference wenbo_engine/circuit/import_qiskit.py:1-38).

Duck-typed like the reference: anything with ``num_qubits``, ``data`` (instructions with
``operation.name`` / ``operation.params`` / ``qubits``) and ``find_bit(q).index`` works, so qiskit
itself is not imported.  Same supported basis, same error text."""
from __future__ import annotations

SUPPORTED_BASIS = ["h", "x", "y", "z", "s", "t", "ry", "cx", "cz", "swap", "cy"]

_QISKIT_MAP = {"h": "H", "x": "X", "y": "Y", "z": "Z", "s": "S", "t": "T", "ry": "RY",
               "cx": "CNOT", "cnot": "CNOT", "swap": "SWAP", "cz": "CZ", "cy": "CY"}
_SKIP = frozenset({"barrier", "measure", "reset", "delay", "id"})


def qiskit_to_dict(qc) -> dict:
    """Convert an (already transpiled) QuantumCircuit; barriers / measurements are dropped."""
    gates = []
    for inst in qc.data:
        op = inst.operation
        name = op.name.lower()
        if name in _SKIP:
            continue
        if name not in _QISKIT_MAP:
            raise ValueError(f"Unsupported gate '{name}'. Transpile to basis {SUPPORTED_BASIS} first.")
        entry: dict = {"qubits": [qc.find_bit(q).index for q in inst.qubits], "gate": _QISKIT_MAP[name], "params": {}}
        if name == "ry":
            entry["params"]["theta"] = float(op.params[0])
        gates.append(entry)
    return {"number_of_qubits": qc.num_qubits, "gates": gates}
