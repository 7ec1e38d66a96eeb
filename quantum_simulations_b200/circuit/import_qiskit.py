"""Front end for Qiskit circuits: ``qiskit_to_dict(qc)`` gives the circuit dict that
``runner.single_node.run`` / ``kernel.cuda_dense.simulate`` consume.

Behavioural mirror of the reference's importer (wenbo_engine/circuit/import_qiskit.py): the same
eleven basis gates are accepted, scheduling / measurement instructions are dropped, anything else
raises ``ValueError("Unsupported gate ...")`` asking for a transpile to SUPPORTED_BASIS.  Nothing of
qiskit is imported: any object that quacks like a transpiled QuantumCircuit is accepted, i.e. it has
``num_qubits``, ``find_bit(bit).index`` and ``data`` whose items carry ``operation`` (with ``name``
and ``params``) and ``qubits``."""
from __future__ import annotations

from typing import Any, Callable

# qiskit name -> (circuit-dict gate name, parameter extractor or None)
_TABLE: dict[str, tuple[str, Callable[[Any], dict] | None]] = {
    "h": ("H", None),
    "x": ("X", None),
    "y": ("Y", None),
    "z": ("Z", None),
    "s": ("S", None),
    "t": ("T", None),
    "ry": ("RY", lambda operation: {"theta": float(operation.params[0])}),
    "cx": ("CNOT", None),
    "cnot": ("CNOT", None),
    "cz": ("CZ", None),
    "swap": ("SWAP", None),
    "cy": ("CY", None),
}
# instructions that carry no unitary action on the state the engine returns
_NOT_GATES = ("barrier", "measure", "reset", "delay", "id")

SUPPORTED_BASIS = ["h", "x", "y", "z", "s", "t", "ry", "cx", "cz", "swap", "cy"]


def _convert(qc, instruction) -> dict | None:
    operation = instruction.operation
    key = str(operation.name).lower()
    if key in _NOT_GATES:
        return None
    if key not in _TABLE:
        raise ValueError(f"Unsupported gate '{key}'. Transpile to basis {SUPPORTED_BASIS} first.")
    gate_name, extract = _TABLE[key]
    return {
        "qubits": [qc.find_bit(bit).index for bit in instruction.qubits],
        "gate": gate_name,
        "params": extract(operation) if extract else {},
    }


def qiskit_to_dict(qc) -> dict:
    """Circuit dict of an (already transpiled) QuantumCircuit, gates in program order."""
    converted = (_convert(qc, instruction) for instruction in qc.data)
    return {"number_of_qubits": int(qc.num_qubits), "gates": [g for g in converted if g is not None]}
