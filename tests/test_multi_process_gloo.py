"""N > 1 host logic on CPU: world_size 2 and 4, one process per shard, gloo backend."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from oracle import ref_dense as O
from quantum_simulations_b200 import workloads as W
from quantum_simulations_b200.circuit.io import validate_circuit_dict

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_run_across_processes(tmp_path, world):
    n = 10
    port = 29600 + world + (os.getpid() % 200)
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(ROOT / "tests" / "gloo_worker.py"), str(tmp_path), str(n)],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    for name, cd in (("random_1q_cz", W.random_1q_cz(n, 12, 99)), ("qft", W.qft(n)), ("random_mixed", W.random_mixed(n, 100, 4))):
        got = np.concatenate([np.load(tmp_path / f"{name}_rank{r}.npy") for r in range(world)])
        want = O.simulate(validate_circuit_dict(cd))
        assert np.abs(got - want).max() <= 1e-12, name
        assert int((tmp_path / f"{name}_swaps.txt").read_text()) >= 1, name
