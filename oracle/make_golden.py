"""Generate tests/golden/*.npz by running the REAL reference (pure Python + NumPy).

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python oracle/make_golden.py

Everything written here is an OUTPUT OF THE UNMODIFIED REFERENCE
(``wenbo_engine`` imported from /root/reference), keyed by a workload spec string that
``tests/_specs.py`` turns back into the same circuit dict through our own generators.
The vectors pin (a) the oracle restatement and (b) the CUDA path.
"""
from __future__ import annotations

import json
import sys
import tempfile
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, "/root/reference")

from wenbo_engine.kernel import gates as rgates            # noqa: E402
from wenbo_engine.kernel import ref_dense as rref           # noqa: E402
from wenbo_engine.kernel import cpu_scalar, cpu_batched, cpu_nonlocal  # noqa: E402
from wenbo_engine.circuit.io import validate_circuit_dict, levelize   # noqa: E402
from wenbo_engine.circuit.fusion import fuse_1q_ops, batch_levels      # noqa: E402
from wenbo_engine.circuit.staging import permute_state                 # noqa: E402
from wenbo_engine.runner import single_node as rsn                     # noqa: E402
from wenbo_engine.wal.wal import _circuit_hash                         # noqa: E402

from tests._specs import circuit_from_spec, STATE_SPECS, RUNNER_SPECS  # noqa: E402

OUT = REPO / "tests" / "golden"
OUT.mkdir(parents=True, exist_ok=True)


def gate_vectors() -> dict:
    th = 1.2345
    u = np.array([[0.6, -0.8j], [0.8j, 0.6]])
    cases = {
        "H": {}, "X": {}, "Y": {}, "Z": {}, "S": {}, "T": {},
        "RY": {"theta": th}, "R": {"k": 3}, "G": {"p": 5},
        "CNOT": {}, "SWAP": {}, "CZ": {}, "CY": {},
        "CR": {"k": 4}, "CU": {"U": u, "exponent": 3},
    }
    return {f"gate/{k}": rgates.gate_matrix(k, p) for k, p in cases.items()}


def kernel_vectors() -> dict:
    rng = np.random.default_rng(20240)
    out = {}
    n = 8
    chunk = (rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)).astype(np.complex128)
    out["kernel/input"] = chunk
    U1 = rgates.gate_matrix("RY", {"theta": 0.77}) @ rgates.T() @ rgates.H()
    U2 = rgates.gate_matrix("CU", {"U": np.array([[0.6, -0.8j], [0.8j, 0.6]]), "exponent": 1}) \
        @ np.kron(rgates.H(), rgates.T()) @ rgates.SWAP()
    out["kernel/U1"], out["kernel/U2"] = U1, U2
    for q in (0, 1, 4, 7):
        for name, mod in (("scalar", cpu_scalar), ("batched", cpu_batched)):
            c = chunk.copy()
            mod.apply_1q(c, q, U1)
            out[f"kernel/{name}_1q_q{q}"] = c
    for qa, qb in ((0, 1), (1, 0), (2, 6), (7, 3), (7, 0)):
        for name, mod in (("scalar", cpu_scalar), ("batched", cpu_batched)):
            c = chunk.copy()
            mod.apply_2q(c, qa, qb, U2)
            out[f"kernel/{name}_2q_{qa}_{qb}"] = c
    # non-local butterflies on 4 chunks of 64
    parts = [chunk[i * 64:(i + 1) * 64].copy() for i in range(4)]
    a, b = parts[0].copy(), parts[1].copy()
    cpu_nonlocal.apply_1q_pair(a, b, U1)
    out["nonlocal/1q_pair"] = np.concatenate([a, b])
    a, b = parts[0].copy(), parts[1].copy()
    cpu_nonlocal.apply_2q_pair_qa_local(a, b, 3, U2)
    out["nonlocal/qa_local_3"] = np.concatenate([a, b])
    a, b = parts[0].copy(), parts[1].copy()
    cpu_nonlocal.apply_2q_pair_qb_local(a, b, 2, U2)
    out["nonlocal/qb_local_2"] = np.concatenate([a, b])
    q4 = [p.copy() for p in parts]
    cpu_nonlocal.apply_2q_quad(*q4, U2)
    out["nonlocal/quad"] = np.concatenate(q4)
    # permute_state
    st = chunk[:32].copy()
    out["permute/input"] = st
    out["permute/l2p_2_0_1_4_3"] = permute_state(st, [2, 0, 1, 4, 3])
    return out


def host_vectors() -> dict:
    """Levelization / fusion results as JSON-able structures."""
    meta = {}
    for spec in ("qft:5", "ghz:6", "random_1q_cz:7:6:5", "random_mixed:6:40:11"):
        cd = validate_circuit_dict(circuit_from_spec(spec))
        lv = levelize(cd)
        ids = {id(g): i for i, g in enumerate(cd["gates"])}
        meta[f"levelize/{spec}"] = [[ids[id(g)] for g in level] for level in lv]
        steps = batch_levels(lv, 4)
        meta[f"batch_levels_k4/{spec}"] = [
            {"n_local": len(s["local_ops"]), "n_nonlocal": len(s["nonlocal_ops"]),
             "level_indices": s["level_indices"],
             "local_qubits": [list(q) for q, _ in s["local_ops"]]}
            for s in steps]
        meta[f"circuit_hash/{spec}"] = _circuit_hash(cd)
    return meta


def main() -> None:
    arrays = {}
    arrays.update(gate_vectors())
    arrays.update(kernel_vectors())
    for spec in STATE_SPECS:
        arrays[f"state/{spec}"] = rref.simulate(circuit_from_spec(spec))
    np.savez_compressed(OUT / "reference_vectors.npz", **arrays)

    # fusion algebra: reference fuse_1q_ops on a fixed op list
    ops = [([0], rgates.H()), ([1], rgates.T()), ([0], rgates.T()), ([0, 1], rgates.CNOT()),
           ([1], rgates.S()), ([1], rgates.H()), ([2], rgates.X()), ([0], rgates.Y())]
    fused = fuse_1q_ops(ops)
    np.savez_compressed(OUT / "fusion_vectors.npz",
                        **{f"fused/{i}/q{'_'.join(map(str, q))}": U for i, (q, U) in enumerate(fused)})

    # reference out-of-core runner (complex64 storage) on tiny chunks
    runner = {}
    for spec, chunk_size, kw in RUNNER_SPECS:
        with tempfile.TemporaryDirectory() as td:
            final = rsn.run(circuit_from_spec(spec), td, chunk_size=chunk_size, use_wal=False, **kw)
            key = f"runner/{spec}/cs{chunk_size}/" + ",".join(f"{k}={v}" for k, v in sorted(kw.items()))
            runner[key] = rsn.collect_state(final, apply_permutation=True, work_dir=td)
    np.savez_compressed(OUT / "runner_vectors.npz", **runner)

    (OUT / "host_vectors.json").write_text(json.dumps(host_vectors(), indent=1))
    total = sum(f.stat().st_size for f in OUT.iterdir())
    print(f"wrote {len(arrays)} arrays, {len(runner)} runner states; {total / 1e6:.2f} MB in {OUT}")


if __name__ == "__main__":
    main()
