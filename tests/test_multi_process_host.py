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
        procs.append(subprocess.Popen([sys.executable, str(ROOT / "tests" / "multi_process_worker.py"), str(tmp_path), str(n)],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    for name, cd in (("random_1q_cz", W.random_1q_cz(n, 12, 99)), ("qft", W.qft(n)), ("random_mixed", W.random_mixed(n, 100, 4))):
        got = np.concatenate([np.load(tmp_path / f"{name}_rank{r}.npy") for r in range(world)])
        want = O.simulate(validate_circuit_dict(cd))
        assert np.abs(got - want).max() <= 1e-12, name
        assert int((tmp_path / f"{name}_swaps.txt").read_text()) >= 1, name
    _check_qasm(tmp_path, world, n)


def _check_qasm(tmp_path, world, n):
    sys.path.insert(0, str(ROOT / "tests"))
    from multi_process_worker import qasm_text
    from quantum_simulations_b200.circuit.qasm import qasm_to_ops
    _, ops = qasm_to_ops(qasm_text(n))
    want = np.zeros(1 << n, dtype=np.complex128)
    want[0] = 1
    O.apply_ops(want, ops)
    got = np.concatenate([np.load(tmp_path / f"qasm_rank{r}.npy") for r in range(world)])
    assert np.abs(got - want).max() <= 1e-12


def test_bench_multi_control_flow_on_cpu(tmp_path):
    """bench.py --gpus 2 with the GPU objects replaced by emulator fakes (tests/bench_multi_fake.py):
    the N > 1 code path runs end to end over gloo and prints one well-formed JSON line."""
    import json
    world = 2
    port = 29800 + (os.getpid() % 150)
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(ROOT / "tests" / "bench_multi_fake.py"), str(world)],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(o[1][-2000:] for o in outs)
    line = json.loads(outs[0][0].strip().splitlines()[-1])
    assert line["n_gpus"] == 2 and line["scaling"] == "weak" and line["config"]["n_qubits"] == 12
    assert line["config"]["swaps_per_step"] >= 1 and line["config"]["plan"] == "default planner options"
    assert line["e2e"]["value"] > 0 and line["roofline"]["bound"] == "hbm" and "nvlink" in line
    assert outs[1][0].strip() == ""                                  # only rank 0 prints


@pytest.mark.parametrize("broken", [None, "pairs", "search"])
def test_bench_multi_checks_parity_before_it_times(broken):
    """bench.py --gpus 2: the cross-G parity check runs BEFORE the timed region; when it fails under the default
    data-movement switches of the pass kernels the run falls back to QSV_JIT_PAIR=0 / QSV_JIT_TILE_BLOCK=0 on every
    rank, checks again and says so in config.preflight (emulator fakes, real plumbing)."""
    import json
    world = 2
    port = 30400 + (os.getpid() % 150) + {None: 0, "pairs": 37, "search": 74}[broken]
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        for k in ("QSV_JIT_PAIR", "QSV_JIT_TILE_BLOCK", "QSV_PLAN_SEARCH"):
            env.pop(k, None)
        if broken:
            env["FAKE_BREAK_PAIRED_LOADS" if broken == "pairs" else "FAKE_BREAK_PLAN_SEARCH"] = "1"
        procs.append(subprocess.Popen([sys.executable, str(ROOT / "tests" / "bench_multi_fake.py"), str(world), "parity"],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(o[1][-2000:] for o in outs)
    line = json.loads(outs[0][0].strip().splitlines()[-1])
    assert line["parity"]["ok"] is True and line["parity"]["all_indices_covered"] is True
    assert line["parallel_efficiency"]["one_gpu"]["n_qubits"] == 12
    if broken:
        pre = line["config"]["preflight"]
        assert pre["what"].startswith("FALLBACK") and pre["first_check"]["ok"] is False
        assert pre["first_check"]["max_abs_diff"] > 1e-4
        # the most recent default is switched off first; the planner's search only if that did not help
        assert len(pre["levels"]) == (1 if broken == "pairs" else 2)
        assert ("QSV_PLAN_SEARCH=0" in pre["what"]) == (broken == "search") and "QSV_JIT_PAIR=0" in pre["what"]
    else:
        assert line["config"]["preflight"] is None


def test_host_plumbing_collectives(tmp_path):
    """runner/plumbing.py over three real processes: gather, broadcast, AND, barrier, rank-ordered reductions."""
    world = 3
    port = 30100 + (os.getpid() % 200)
    code = (
        "import sys, json, numpy as np\n"
        f"sys.path.insert(0, {str(ROOT)!r})\n"
        "from quantum_simulations_b200.runner.plumbing import init_plumbing, shutdown_plumbing\n"
        "p = init_plumbing(); r = p.rank\n"
        "out = {'gather': p.all_gather_object({'r': r, 'b': bytes([r]) * 4}), 'bcast': p.broadcast_object('x' * (r + 1), src=1),\n"
        "       'all_t': p.all(True), 'all_f': p.all(r != 2), 'sum': p.allreduce(0.1 * (r + 1), 'sum'), 'max': p.allreduce(float(r), 'max'),\n"
        "       'vec': p.allreduce(np.arange(4.0) * (r + 1), 'sum').tolist(), 'big': int(p.allreduce(np.ones(1 << 20), 'sum')[-1])}\n"
        "p.barrier(); shutdown_plumbing()\n"
        "out['gather'] = [[g['r'], g['b'].hex()] for g in out['gather']]\n"
        "print(json.dumps(out))\n")
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=120) for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(o[1][-1500:] for o in outs)
    import json
    res = [json.loads(o[0].strip().splitlines()[-1]) for o in outs]
    assert all(r == res[0] for r in res)                     # every rank sees the same results, bit for bit
    r0 = res[0]
    assert r0["gather"] == [[0, "00000000"], [1, "01010101"], [2, "02020202"]] and r0["bcast"] == "xx"
    assert r0["all_t"] is True and r0["all_f"] is False and r0["max"] == 2.0 and r0["big"] == 3
    assert r0["sum"] == (0.1 + 0.2) + 0.30000000000000004 and r0["vec"] == [0.0, 6.0, 12.0, 18.0]


def test_the_product_imports_no_tensor_library():
    """North star: host side = Python + ctypes over the C ABI; torch may launch the processes (torchrun) but
    nothing under quantum_simulations_b200/ imports it."""
    import re
    bad = []
    for f in (ROOT / "quantum_simulations_b200").rglob("*.py"):
        for i, line in enumerate(f.read_text().splitlines(), 1):
            if re.match(r"\s*(import|from)\s+(torch|triton|jax|cupy)\b", line):
                bad.append(f"{f.relative_to(ROOT)}:{i}: {line.strip()}")
    assert not bad, bad


def test_plumbing_codec_is_closed_and_round_trips():
    """The collectives carry a closed set of types (runner/plumbing.py encode / decode): no pickle on the wire, so a
    process that reaches rank 0's port cannot make it execute anything; malformed messages are refused."""
    from quantum_simulations_b200.runner import plumbing as P
    assert "pickle" not in Path(P.__file__).read_text().split('"""', 2)[2]          # (the docstring explains why)
    vals = [None, True, False, 0, -1, 2 ** 70, -(2 ** 65), 1.5, float("inf"), "héllo", b"\x00\x01", [1, [2, (3, None)]], (), {},
            {1: (2, b"x"), "a": [1.0]}, np.arange(6.0).reshape(2, 3), np.array(3.0), np.zeros((0,), dtype=np.int64),
            np.arange(4) * 1j, np.float64(2.5), np.int64(7), np.bool_(True), np.uint64(2 ** 63)]
    for v in vals:
        w = P.decode(bytes(P.encode(v)))
        if isinstance(v, np.ndarray):
            assert w.dtype == v.dtype and w.shape == v.shape and np.array_equal(w, v)
        else:
            assert w == v and not isinstance(w, np.generic)
    both = P.decode(bytes(P.encode((np.arange(3), np.arange(3) * 1j))))
    assert isinstance(both, tuple) and np.array_equal(both[1], np.arange(3) * 1j)
    for bad in (b"", b"Z", b"I\x05\x00\x00\x00ab", b"L" + b"\xff" * 8, b"NN", b"S\x02" + b"\x00" * 7 + b"a",
                b"A\x03\x01<f8" + b"\x02" + b"\x00" * 7 + b"\x08" + b"\x00" * 7 + b"\x00" * 8,        # shape (2,) but 8 bytes
                b"A\x02\x00|O" + b"\x08" + b"\x00" * 7 + b"\x00" * 8,                                  # object dtype
                b"A\x03\x00xyz" + b"\x00" * 8):                                                        # no such dtype
        with pytest.raises(ValueError):
            P.decode(bad)
    for refused in (object(), np.array([object()]), {object(): 1}, lambda: 0):
        with pytest.raises(TypeError):
            P.encode(refused)
