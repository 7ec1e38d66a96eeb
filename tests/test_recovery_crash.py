"""WAL + crash recovery of the GPU runner — the reference's tests/test_recovery_crash.py restated
for kernel="cuda" (device-resident state, checkpoints through pinned async copies)."""
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from oracle import ref_dense as O
from quantum_simulations_b200 import workloads as W
from quantum_simulations_b200.circuit.io import validate_circuit_dict
from quantum_simulations_b200.wal.fencing import FencingLock
from quantum_simulations_b200.wal.recovery import recover
from quantum_simulations_b200.wal.wal import WAL

ROOT = Path(__file__).resolve().parent.parent


def _bell():
    return {"number_of_qubits": 2, "gates": [{"qubits": [0], "gate": "H"}, {"qubits": [0, 1], "gate": "CNOT"}]}


# ---------------------------------------------------------------- host-only pieces
def test_no_wal_no_recovery(tmp_path):
    assert recover(_bell(), tmp_path) is None


def test_fencing_lock(tmp_path):
    lock = FencingLock(tmp_path)
    lock.acquire()
    assert (tmp_path / "run.lock").exists()
    with pytest.raises(RuntimeError, match="locked by PID"):
        FencingLock(tmp_path).acquire()
    lock.release()
    assert not (tmp_path / "run.lock").exists()
    with FencingLock(tmp_path):
        assert (tmp_path / "run.lock").exists()
    assert not (tmp_path / "run.lock").exists()


def test_wal_circuit_hash_mismatch(tmp_path):
    WAL(tmp_path / "wal.json", circuit_dict=_bell()).close()
    with pytest.raises(ValueError, match="circuit hash mismatch"):
        WAL(tmp_path / "wal.json", circuit_dict={"number_of_qubits": 2, "gates": [{"qubits": [0], "gate": "X"}]})


def test_qiskit_import_mirror():
    from types import SimpleNamespace as NS
    from quantum_simulations_b200.circuit.import_qiskit import qiskit_to_dict

    class QC:
        num_qubits = 3
        data = [NS(operation=NS(name="h", params=[]), qubits=["q0"]),
                NS(operation=NS(name="barrier", params=[]), qubits=["q0", "q1"]),
                NS(operation=NS(name="ry", params=[0.25]), qubits=["q2"]),
                NS(operation=NS(name="cx", params=[]), qubits=["q0", "q2"])]

        @staticmethod
        def find_bit(q):
            return NS(index=int(q[1:]))

    cd = qiskit_to_dict(QC())
    assert cd == {"number_of_qubits": 3, "gates": [
        {"qubits": [0], "gate": "H", "params": {}}, {"qubits": [2], "gate": "RY", "params": {"theta": 0.25}},
        {"qubits": [0, 2], "gate": "CNOT", "params": {}}]}
    QC.data.append(NS(operation=NS(name="rz", params=[1.0]), qubits=["q0"]))
    with pytest.raises(ValueError, match="Unsupported gate 'rz'"):
        qiskit_to_dict(QC())


# ---------------------------------------------------------------- on the device
_SCRIPT = """
import sys, json
sys.path.insert(0, {root!r})
from quantum_simulations_b200.runner.{runner} import run
cd = json.loads({cd!r})
run(cd, {work!r}, chunk_size={cs}, use_wal=True, checkpoint_every=1)
"""


def _run_sub(cd, work, cs, crash_after=None, at_checkpoint=0, runner="single_node"):
    env = os.environ.copy()
    env["WE_CRASH_AT_CHECKPOINT"] = str(at_checkpoint)
    if crash_after is not None:
        env["WE_CRASH_AFTER_CHUNK"] = str(crash_after)
    else:
        env.pop("WE_CRASH_AFTER_CHUNK", None)
    script = _SCRIPT.format(root=str(ROOT), cd=json.dumps(cd), work=str(work), cs=cs, runner=runner)
    return subprocess.run([sys.executable, "-c", script], env=env, capture_output=True, timeout=300)


@pytest.mark.gpu
def test_runner_wal_commits_and_alternates(tmp_path):
    from quantum_simulations_b200.runner.single_node import run, collect_state, _buf_dir
    cd = _bell()
    final = run(cd, tmp_path, chunk_size=4, use_wal=True, checkpoint_every=1)
    wal = WAL(tmp_path / "wal.json", circuit_dict=cd)
    assert wal.done_steps == 2 and wal.committed_buf == "a"          # a -> b -> a (reference test 7)
    assert (_buf_dir(tmp_path, wal.committed_buf) / "manifest.json").exists()
    assert np.abs(collect_state(final) - O.simulate(validate_circuit_dict(cd))).max() <= 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("crash_after", [0, 1, 3])
def test_crash_env_var_and_recover(tmp_path, crash_after):
    """Die inside a checkpoint (WE_CRASH_AFTER_CHUNK, reference single_node.py:61-63); the committed
    buffer stays intact, a second run resumes from wal.done_steps and ends at the oracle's state."""
    from quantum_simulations_b200.runner.single_node import collect_state
    n = 12
    cd = W.random_1q_cz(n, 6, 5)                                     # 6 levels -> 6 checkpoints of 4 chunks
    cs = 1 << (n - 2)
    r = _run_sub(cd, tmp_path, cs, crash_after)
    assert r.returncode != 0, r.stderr.decode()[-500:]
    wal = WAL(tmp_path / "wal.json", circuit_dict=cd)
    assert wal.done_steps == 0                                        # died before the first commit
    (tmp_path / "state_b" / "chunks").mkdir(parents=True, exist_ok=True)
    (tmp_path / "state_b" / "chunks" / "garbage.bin").write_bytes(b"\\xff" * 100)
    final = recover(cd, tmp_path, chunk_size=cs, checkpoint_every=1)
    assert final is not None
    assert np.abs(collect_state(final) - O.simulate(validate_circuit_dict(cd))).max() <= 1e-12
    assert not (tmp_path / "state_b" / "chunks" / "garbage.bin").exists()


@pytest.mark.gpu
def test_resume_from_a_committed_checkpoint(tmp_path):
    """Crash in the THIRD checkpoint: two steps are committed, the rerun reloads that checkpoint
    into HBM (it does not start over) and finishes."""
    from quantum_simulations_b200.runner.single_node import collect_state
    n = 12
    cd = W.random_1q_cz(n, 6, 9)
    cs = 1 << (n - 2)
    r = _run_sub(cd, tmp_path, cs, crash_after=2, at_checkpoint=2)
    assert r.returncode != 0, r.stderr.decode()[-500:]
    wal = WAL(tmp_path / "wal.json", circuit_dict=cd)
    assert wal.done_steps == 2 and wal.committed_buf == "a"
    final = recover(cd, tmp_path, chunk_size=cs, checkpoint_every=1)
    assert np.abs(collect_state(final) - O.simulate(validate_circuit_dict(cd))).max() <= 1e-12


@pytest.mark.gpu
def test_pipeline_runner_checkpoints_asynchronously_and_is_recoverable(tmp_path):
    """runner.pipeline.run: checkpoints drain from a device snapshot on a writer thread while the next steps run.
    Same files, WAL and crash behaviour as single_node: (1) a full run equals the oracle and commits every step;
    (2) a crash inside the third checkpoint leaves two committed steps, and single_node's recover() finishes the
    directory the pipeline runner wrote."""
    from quantum_simulations_b200.runner import pipeline
    from quantum_simulations_b200.runner.single_node import collect_state
    n = 12
    cd = W.random_1q_cz(n, 6, 9)
    cs = 1 << (n - 3)
    want = O.simulate(validate_circuit_dict(cd))
    final = pipeline.run(cd, tmp_path / "full", chunk_size=cs, buffer_depth=3, use_wal=True, checkpoint_every=1)
    wal = WAL(tmp_path / "full" / "wal.json", circuit_dict=cd)
    assert wal.done_steps == 6 and np.abs(collect_state(final) - want).max() <= 1e-12
    final = pipeline.run(cd, tmp_path / "sparse", chunk_size=cs, buffer_depth=2, use_wal=True, checkpoint_every=4, use_fusion=False)
    assert WAL(tmp_path / "sparse" / "wal.json", circuit_dict=cd).done_steps == 6
    assert np.abs(collect_state(final) - want).max() <= 1e-12
    work = tmp_path / "crash"
    r = _run_sub(cd, work, cs, crash_after=2, at_checkpoint=2, runner="pipeline")
    assert r.returncode != 0, r.stderr.decode()[-500:]
    wal = WAL(work / "wal.json", circuit_dict=cd)
    assert wal.done_steps == 2 and wal.committed_buf == "a"
    final = recover(cd, work, chunk_size=cs, checkpoint_every=1)
    assert np.abs(collect_state(final) - want).max() <= 1e-12
