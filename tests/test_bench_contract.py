"""The bench line the driver parses: checked on the line committed under profiles/ (produced by
``python bench.py`` on a B200) and on the argument parser, so a refactor cannot silently drop a key."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _line(name):
    return json.loads((ROOT / "profiles" / "r01" / name).read_text().strip().splitlines()[-1])


def test_single_gpu_line_has_every_contract_key():
    r = _line("bench_n30_jit.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in r, k
    assert r["metric"] == "amplitude-updates/s" and r["n_gpus"] == 1 and r["higher_is_better"] is True
    assert r["dtype"] == "complex128" and r["data"] == "synthetic" and r["vs_baseline"] is None
    assert "workload" in r["config"] and "model" not in r["config"] and r["config"]["n_qubits"] == 30
    rf = r["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert 0 < rf["frac"] <= 1.0 and rf["traffic"] and abs(rf["traffic"] / rf["algorithmic_bytes_per_launch"] - 1) < 0.01
    cb = r["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] > 0
    e = r["e2e"]
    assert e["value"] > 0 and e["d2h_bytes_per_step"] == 16 * 2 ** 30 and e["h2d_bytes_per_step"] > 0
    assert e["value"] < r["value"]                       # end to end includes the copies
    assert r["gpu_launches"] > 0
    assert set(r["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert abs(r["value"] - r["config"]["gates"] * 2 ** 30 / (r["ms_per_step"] * 1e-3)) / r["value"] < 1e-6


def test_multi_gpu_lines():
    for name, n_gpus in (("bench_n31_2gpu_peer_swap.json", 2), ("bench_n32_4gpu.json", 4), ("bench_n33_8gpu.json", 8),
                         ("bench_n36_8gpu_1TiB.json", 8)):
        r = _line(name)
        assert r["n_gpus"] == n_gpus and r["scaling"] == "weak" and r["metric"] == "amplitude-updates/s"
        assert r["roofline"]["bound"] == "hbm" and "nvlink" in r and r["config"]["swaps_per_step"] >= 1


def test_cli_surface():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--help"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl"):
        assert flag in out.stdout


def test_single_gpu_control_flow_on_cpu():
    """bench.py at N = 1 with the GPU objects replaced by emulator fakes (tests/bench_single_fake.py):
    the code path the driver runs assembles a well-formed line (incl. roofline without the init pass)."""
    out = subprocess.run([sys.executable, str(ROOT / "tests" / "bench_single_fake.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert r["n_gpus"] == 1 and r["config"]["n_qubits"] == 12 and r["e2e"]["value"] > 0 and r["cpu_baseline"]["value"] > 0
    assert r["config"]["init_note"] is None and r["config"]["init_pass_ms"] is not None
    assert r["roofline"]["launches_timed"] == (r["config"]["passes_per_step"] - 1) * r["steps"]
    assert r["zero_support_skipping"]["ms_per_step"] > 0
    assert r["full_first_pass"]["ms_per_step"] > 0 and r["config"]["planner_switches"]["low_store_round"] is True


def _line2(name):
    return json.loads((ROOT / "profiles" / "r02" / name).read_text().strip().splitlines()[-1])


def test_round2_single_gpu_line():
    r = _line2("bench_n30_default.json")
    assert r["n_gpus"] == 1 and r["config"]["n_qubits"] == 30 and r["dtype"] == "complex128"
    assert r["cpu_baseline"]["kind"] == "reference"                     # the unmodified reference, from oracle/_ref
    rf = r["roofline"]
    assert 0.8 < rf["frac"] <= 1.0 and abs(rf["traffic"] / rf["algorithmic_bytes_per_launch"] - 1) < 0.01
    assert r["e2e"]["cold"]["jit"]["kernels_compiled"] >= 1 and r["e2e"]["cold"]["ms_per_step"] > r["e2e"]["ms_per_step"]
    names = [(o["n_qubits"], o["dtype"]) for o in r["other_workloads"]]
    assert (28, "complex128") in names and (30, "complex64") in names and (20, "complex128") in names
    assert all(abs(o["norm"] - 1) < 1e-6 for o in r["other_workloads"])


def test_round2_multi_gpu_lines_are_the_baseline_configs():
    """bench.py --gpus N runs BASELINE configs[3] / [4] by default and carries parity, efficiency and the weak series."""
    for name, n_gpus, n in (("bench_n34_2gpu_final.json", 2, 34), ("bench_n34_4gpu_final.json", 4, 34),
                            ("bench_n36_8gpu_1TiB_final.json", 8, 36)):
        r = _line2(name)
        assert r["n_gpus"] == n_gpus and r["config"]["n_qubits"] == n and r["metric"] == "amplitude-updates/s"
        assert r["parity"]["ok"] and r["parity"]["sampled_amplitudes"] >= 1 << 20 and r["parity"]["max_abs_diff"] <= 1e-12
        assert r["pipelined_swap"]["count_per_step"] >= 1 and r["nvlink"]["share_of_step"] < 0.15
        pe = r["parallel_efficiency"]
        assert abs(pe["value"] - r["value"] / (n_gpus * pe["one_gpu"]["value"])) < 1e-9 and pe["value"] >= 0.60
        assert r["weak_series"]["n_qubits"] == 30 + n_gpus.bit_length() - 1
        assert r["e2e"]["d2h_bytes_per_step"] == 16 * 2 ** n
