"""Generate tests/golden/v1_sqlite_states.json by running the reference's FIRST implementation — the SQLite
amplitude-table simulator (v1_implementation/src: gates as SQL JOIN + GROUP BY over rows (idx, real, imag),
gate_translator.py:9-55, simulator.py:19-29) — unmodified, on an in-memory database.

Run in the build container only (/root/reference does not exist on the GPU box):

    python oracle/make_golden_v1.py

It is an implementation INDEPENDENT of wenbo_engine (sparse rows, SQL arithmetic), so its outputs pin the oracle
(and through it the CUDA path) a second time, and they pin the (idx, real, imag) row format that
storage/sparse_rows.py imports and exports.  Circuits come from v1's own
generators (src/circuits.py); each is stored as the circuit dict it ran (matrices of CU gates as nested lists)
together with the rows v1 produced."""
from __future__ import annotations

import json
import sqlite3
import sys
import tempfile
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(REF))

from v1_implementation.src import circuits as C1                       # noqa: E402
from v1_implementation.src import db as DB1                            # noqa: E402
from v1_implementation.src.simulator import run_circuit               # noqa: E402
from v1_implementation.src.state_manager import fetch_state           # noqa: E402


def jsonable(cd: dict) -> dict:
    gates = []
    for g in cd["gates"]:
        p = dict(g.get("params", {}))
        if "U" in p:
            u = np.asarray(p["U"], dtype=np.complex128)
            p["U"] = {"re": u.real.tolist(), "im": u.imag.tolist()}
        gates.append({"qubits": list(g["qubits"]), "gate": g["gate"], "params": p})
    return {"number_of_qubits": cd["number_of_qubits"], "gates": gates}


def main() -> None:
    cases = {
        "ghz5": C1.generate_ghz_circuit(5),
        "qft4": C1.generate_qft_circuit(4),
        "qft6": C1.generate_qft_circuit(6),
        "qpe3": C1.generate_qpe_circuit(3),
        "w5": C1.generate_w_circuit(5),
        "hwall4": C1.generate_hadamard_wall(4),
        "w_qft4": C1.generate_w_qft(4),
        "ghz_qft5": C1.generate_ghz_qft(5),
        "mixed4": {"number_of_qubits": 4, "gates": [
            {"qubits": [0], "gate": "H"}, {"qubits": [1], "gate": "RY", "params": {"theta": 0.7}},
            {"qubits": [2], "gate": "T"}, {"qubits": [3], "gate": "Y"}, {"qubits": [0, 3], "gate": "CZ"},
            {"qubits": [1, 2], "gate": "CY"}, {"qubits": [3, 1], "gate": "SWAP"}, {"qubits": [2], "gate": "S"},
            {"qubits": [0], "gate": "R", "params": {"k": 3}}, {"qubits": [2, 0], "gate": "CR", "params": {"k": 2}},
            {"qubits": [1], "gate": "G", "params": {"p": 3}}, {"qubits": [3], "gate": "X"}, {"qubits": [1], "gate": "Z"}]},
    }
    out = {}
    schema = REF / "v1_implementation" / "sql" / "schema.sql"
    for name, cd in cases.items():
        with tempfile.TemporaryDirectory() as td:
            con = sqlite3.connect(":memory:")
            DB1.initialize_schema(con, schema)
            version = run_circuit(con, cd, checkpoint_dir=td)
            rows = fetch_state(con, version)
            con.close()
        out[name] = {"circuit": jsonable(cd), "rows": [[int(i), float(r), float(im)] for i, r, im in rows]}
        print(f"{name}: n={cd['number_of_qubits']} gates={len(cd['gates'])} rows={len(rows)}")
    (REPO / "tests" / "golden" / "v1_sqlite_states.json").write_text(json.dumps(out, indent=0))


if __name__ == "__main__":
    main()
