#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's workload.

    python bench.py --gpus 1 --steps K --warmup W            # our arm (CUDA, libqsv.so)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU algorithm

metric  amplitude-updates/s = (#gates * 2^n) per circuit execution / time  (SURVEY.md §8d:
        every gate "updates" all amplitudes, the reference's own bench accounting,
        wenbo_engine/bench/kernel.py:21)
step    one execution of the whole circuit from |0...0>: init + every compiled pass.
N=1     configs[2]: random depth-20 (1q+CZ layers) circuit, 30 qubits, complex128, one B200.
value   device-timed (CUDA events on the library's stream), state resident in HBM.
e2e     the public call ``kernel.cuda_dense.simulate(circuit, out=host_buffer)``: compile,
        allocate, run, and copy the full 2^n state back into pinned HOST memory, host-timed.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from quantum_simulations_b200 import workloads as W                      # noqa: E402
from quantum_simulations_b200.circuit.io import levelize, validate_circuit_dict   # noqa: E402

METRIC = "amplitude-updates/s"
UNIT = "amp-updates/s"


def _peaks() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int = 0):
        self.rows: list[list[str]] = []
        self.proc = None
        self.device = device

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


WORKLOAD = "random"      # --workload: random (configs[2..4]) | qft (configs[1]) | ghz (configs[0])


def workload(n: int) -> tuple[dict, dict]:
    if WORKLOAD == "qft":
        cd = validate_circuit_dict(W.qft(n))
        name = f"qft(n={n}): H(j), CR_(k-j+1)(k,j) (wenbo_engine/tests/fixtures/circuits.py:57-63)"
    elif WORKLOAD == "ghz":
        cd = validate_circuit_dict(W.ghz(n))
        name = f"ghz(n={n}): H(0), CNOT(q-1,q) (wenbo_engine/tests/fixtures/circuits.py:50-54)"
    else:
        cd = validate_circuit_dict(W.random_1q_cz(n, 20, 1234))
        name = (f"random_1q_cz(n={n}, depth=20, seed=1234): alternating 1-qubit "
                f"{{H,X,Y,S,T,RY}} and brickwork CZ layers")
    info = {"workload": name, "n_qubits": n, "gates": len(cd["gates"]), "levels": len(levelize(cd))}
    return cd, info


# ----------------------------------------------------------------------- CPU arms
def _reference_simulate():
    """wenbo_engine.kernel.ref_dense.simulate of the UNMODIFIED reference (oracle/_ref/, placed there by
    __graft_entry__.build() from /root/reference; git-ignored, travels with the built files), or None."""
    ref = ROOT / "oracle" / "_ref"
    if not (ref / "wenbo_engine" / "kernel" / "ref_dense.py").is_file():
        return None
    if str(ref) not in sys.path:
        sys.path.insert(0, str(ref))
    try:
        from wenbo_engine.kernel.ref_dense import simulate as ref_simulate
        return ref_simulate
    except Exception:
        return None


def _cpu_rate(cd: dict) -> tuple[float, float, str]:
    """(amp-updates/s, seconds, kind) of the reference's CPU statevector path on `cd`: the reference's own
    ref_dense.simulate when oracle/_ref holds it (kind "reference"), else its restatement oracle/ref_dense.py
    simulate(indexed=True) (kind "port": the same NumPy index-array formulation, ref_dense.py:13-41)."""
    ref = _reference_simulate()
    t0 = time.perf_counter()
    if ref is not None:
        ref(cd)
        kind = "reference"
    else:
        from oracle import ref_dense as O
        O.simulate(cd, indexed=True)
        kind = "port"
    dt = time.perf_counter() - t0
    return len(cd["gates"]) * (1 << cd["number_of_qubits"]) / dt, dt, kind


def _cpu_c_rate(cd: dict) -> tuple[float, float, int]:
    from oracle import c_oracle as CO
    CO.lib()
    t0 = time.perf_counter()
    CO.simulate_c(cd)
    dt = time.perf_counter() - t0
    return len(cd["gates"]) * (1 << cd["number_of_qubits"]) / dt, dt, CO.n_threads()


def _blas_threads() -> int:
    """threads NumPy's BLAS uses here: the reference's 2-qubit gates are a (4 x 4) @ (4 x M) product per gate
    (ref_dense.py:41), which OpenBLAS runs on its thread pool; everything else is single-threaded NumPy"""
    try:
        from threadpoolctl import threadpool_info
        return max([int(i.get("num_threads", 1)) for i in threadpool_info() if i.get("user_api") == "blas"] or [1])
    except Exception:
        return 1


_WHAT = {"reference": "wenbo_engine.kernel.ref_dense.simulate of the unmodified reference (oracle/_ref)",
         "port": "oracle/ref_dense.py simulate(indexed=True) = the reference's NumPy gather/scatter formulation restated"}


def cpu_baseline(n_sample: int = 20) -> dict:
    cd, _ = workload(n_sample)
    rate, dt, kind = _cpu_rate(cd)
    out = {"value": rate, "unit": UNIT, "cores": _blas_threads(), "kind": kind,
           "sample": f"same circuit family at n={n_sample} (random_1q_cz depth 20, {len(cd['gates'])} gates), "
                     f"{_WHAT[kind]}, complex128; cores = the BLAS threads its 2-qubit (4 x 4) @ (4 x M) products may use, "
                     f"the gather / scatter and 1-qubit arithmetic are single-threaded NumPy; {dt:.1f} s",
           "host_cores_available": os.cpu_count()}
    try:
        cdc, _ = workload(min(n_sample + 4, 24))
        crate, cdt, thr = _cpu_c_rate(cdc)
        out["c_openmp_port"] = {"value": crate, "unit": UNIT, "cores": thr,
                                "sample": f"oracle/ref_dense_c.c at n={cdc['number_of_qubits']}, {cdt:.1f} s"}
    except Exception as e:  # the C port is optional context
        out["c_openmp_port"] = {"error": str(e)[:100]}
    return out


def reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to its workers when it is unset, which OpenBLAS honours: give the reference's
    # BLAS calls every host core again, as in a plain `python bench.py --impl reference` (N = 1)
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count() or 1, user_api="blas")
    except Exception:
        pass
    budget = 150.0 / max(args.steps + args.warmup, 1)          # seconds per step
    probe_cd, _ = workload(14)
    rate_guess, _, kind = _cpu_rate(probe_cd)
    rate_guess *= 0.6                                          # the probe is cache resident, the sample is not
    n_sample = 14
    for n in range(14, 25):
        cd, _ = workload(n)
        if len(cd["gates"]) * (1 << n) / rate_guess <= budget:
            n_sample = n
    cd, info = workload(n_sample)
    for _ in range(args.warmup):
        _cpu_rate(cd)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _cpu_rate(cd)
    dt = time.perf_counter() - t0
    value = args.steps * len(cd["gates"]) * (1 << n_sample) / dt
    sample = (f"bounded sample: same circuit family at n={n_sample} ({len(cd['gates'])} gates, complex128), "
              f"{_WHAT[kind]}; cores = the BLAS threads its 2-qubit (4 x 4) @ (4 x M) products may use, the gather / scatter "
              f"and 1-qubit arithmetic are single-threaded NumPy")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "complex128",
            "data": "synthetic", "config": {**info, "note": "CPU arm runs a bounded sample of the GPU arm's workload"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": _blas_threads(), "kind": kind, "sample": sample,
                             "host_cores_available": os.cpu_count()},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def split_init_pass(per_launch, steps: int, fused_init: bool):
    """(streaming pass times, init pass times).  With the initialisation fused into the first pass, that
    launch reads nothing (it zero-fills the shard and computes the one tile that holds amplitude 0): it is
    kept out of the roofline of the streaming pass kernel, whose launches all read and write the state once."""
    per_step = len(per_launch) // max(steps, 1)
    init_at = set(range(0, len(per_launch), per_step)) if fused_init and per_step and per_launch[0][1] == 10 else set()
    streamed = [ms for k, (ms, kind, _) in enumerate(per_launch) if kind == 10 and k not in init_at]
    init = [per_launch[k][0] for k in sorted(init_at)]
    return streamed, init


def jit_stats() -> dict:
    import ctypes as C
    from quantum_simulations_b200 import _lib as L
    v = [C.c_int() for _ in range(4)]
    sec = C.c_double()
    L.load().qsv_jit_stats(*[C.byref(x) for x in v], C.byref(sec))
    return {"kernels_compiled": v[0].value, "disk_cache_hits": v[1].value, "memory_cache_hits": v[2].value,
            "failed": v[3].value, "nvrtc_seconds": round(sec.value, 3)}


# ----------------------------------------------------------------------- GPU arm
def _other_workload(args, wl: str, n: int, dtype: str, steps: int = 5, warmup: int = 3) -> dict:
    """A short device-timed run of another BASELINE workload (configs[1] QFT-28, GHZ, complex64) with the same
    accounting as the headline: value = gates * 2^n / time, roofline fraction of its streaming pass launches."""
    global WORKLOAD
    from quantum_simulations_b200.circuit.sharding import plan_single
    from quantum_simulations_b200.kernel.cuda import DeviceState
    from quantum_simulations_b200.kernel.cuda_dense import circuit_ops
    keep, WORKLOAD = WORKLOAD, wl
    try:
        cd, info = workload(n)
    finally:
        WORKLOAD = keep
    prog = plan_single(circuit_ops(cd), n, dtype, True, False)
    amp_bytes = np.dtype(dtype).itemsize
    with DeviceState(n, dtype, args.device) as st:
        h = st.upload_program(prog)
        for _ in range(warmup):
            st.replay(h)
        st.sync()
        st.timing(True)
        st.timer_start()
        for _ in range(steps):
            st.replay(h)
        ms = st.timer_stop() / steps
        per_launch = st.take_timings()
        st.timing(False)
        norm = st.norm2()
    streamed, _ = split_init_pass(per_launch, steps, prog.fused_init)
    peak, _ = _peaks()
    avg = float(np.mean(streamed)) if streamed else None
    return {"workload": info["workload"], "n_qubits": n, "dtype": dtype, "gates": info["gates"], "levels": info["levels"],
            "passes_per_step": len(prog.passes), "ms_per_step": ms, "value": info["gates"] * (1 << n) / (ms * 1e-3), "unit": UNIT,
            "gate_layers_per_s": info["levels"] / (ms * 1e-3), "norm": norm,
            "roofline_frac": None if not avg else 2 * amp_bytes * (1 << n) / (avg * 1e-3) / 1e9 / peak,
            "avg_pass_ms": avg, "steps": steps, "warmup": warmup}


def bench_single(args) -> None:
    from quantum_simulations_b200.kernel.cuda import DeviceState
    from quantum_simulations_b200.kernel.cuda_dense import circuit_ops, simulate
    from quantum_simulations_b200.circuit.passes import PassCompiler
    from quantum_simulations_b200.storage.pinned import PinnedBuffer

    n, dtype = args.qubits, args.dtype
    cd, info = workload(n)
    amp_bytes = np.dtype(dtype).itemsize
    t0 = time.perf_counter()
    ckw = dict(tile_bits=args.tile_bits, low_bits=args.low_bits, max_rounds=args.max_rounds,
               defer_diagonals=args.defer_diagonals, fold_tables=not args.no_fold_tables)
    if args.no_low_store_round:
        ckw["low_store_round"] = False
    if args.low_store_bits is not None:
        ckw["low_store_bits"] = args.low_store_bits
    if args.warp_local_rounds:
        ckw["warp_local_rounds"] = True
        os.environ["QSV_JIT_WARP_SYNC"] = "1"
    if args.streaming_stores:
        os.environ["QSV_JIT_STCS"] = "1"
    from quantum_simulations_b200.circuit.sharding import plan_single
    prog = plan_single(circuit_ops(cd), n, dtype, True, False, **ckw)      # from |0...0>: free initial placement
    compile_s = time.perf_counter() - t0
    n_pass = len(prog.passes)
    updates_per_step = len(cd["gates"]) * (1 << n)

    def one_step(st_, handle_, prog_):
        if not prog_.fused_init:               # fused: the first pass creates |0...0> itself
            st_.init_zero()
        st_.replay(handle_)

    fingerprints: list = []

    def timed_run():
        with DeviceState(n, dtype, args.device) as st:
            handle = st.upload_program(prog)
            for _ in range(args.warmup):
                one_step(st, handle, prog)
            st.sync()
            clocks = ClockSampler(args.device).start()
            st.timing(True)
            st.timer_start()
            for _ in range(args.steps):
                one_step(st, handle, prog)
            total_ms_ = st.timer_stop()
            per_launch_ = st.take_timings()
            st.timing(False)
            clk_ = clocks.stop()
            norm_ = st.norm2()
            # fingerprint of the final state: the data-movement switches of the pass kernel (tile-to-CTA mapping,
            # paired loads) must not change a single bit of it
            import hashlib
            cnt = min(1 << 14, 1 << n)
            fp_ = hashlib.sha1(st.download(count=cnt).tobytes()
                               + st.download(offset=((1 << n) // 3) & ~(cnt - 1), count=cnt).tobytes()).hexdigest()[:16]
            fingerprints.append(fp_)
            return total_ms_, per_launch_, clk_, norm_

    norm_tol = 1e-9 if dtype == "complex128" else 1e-4
    total_ms, per_launch, clk, norm = timed_run()
    init_note = None
    if abs(norm - 1.0) > norm_tol and prog.fused_init and "QSV_INIT_PASS_FULL" not in os.environ:
        # insurance for the zero-fill + one-tile form of the first pass (written after the last GPU run
        # of round 1): fall back to the full zero-input launch, which is the measured and tested one
        os.environ["QSV_INIT_PASS_FULL"] = "1"
        init_note = f"FALLBACK: QSV_INIT_PASS_FULL=1 (norm was {norm} with the zero-fill form of the first pass)"
        total_ms, per_launch, clk, norm = timed_run()
    if abs(norm - 1.0) > norm_tol:
        raise SystemExit(f"bench: state norm {norm} != 1 — result invalid")

    ms_per_step = total_ms / args.steps
    value = updates_per_step / (ms_per_step * 1e-3)

    # beside the headline: the same run with the first pass launched over EVERY tile (no zero-fill
    # shortcut), i.e. seven full sweeps of the pass kernel — the conservative reading of the step
    full_first = None
    if prog.fused_init and "QSV_INIT_PASS_FULL" not in os.environ:
        os.environ["QSV_INIT_PASS_FULL"] = "1"
        try:
            f_ms, _, _, f_norm = timed_run()
        finally:
            os.environ.pop("QSV_INIT_PASS_FULL", None)
        if abs(f_norm - 1.0) <= norm_tol:
            full_first = {"ms_per_step": f_ms / args.steps, "value": updates_per_step / (f_ms / args.steps * 1e-3), "unit": UNIT,
                          "what": "QSV_INIT_PASS_FULL=1: the zero-input first pass runs its arithmetic on every tile"}

    # the same circuit with ZERO-SUPPORT SKIPPING (compile(zero_state=True)): reported beside the
    # headline, never as the headline — the roofline accounting assumes every pass streams the state
    zs = None
    if not args.no_zero_support:
        prog_z = plan_single(circuit_ops(cd), n, dtype, True, True, **ckw)
        with DeviceState(n, dtype, args.device) as st:
            hz = st.upload_program(prog_z)
            for _ in range(args.warmup):
                one_step(st, hz, prog_z)
            st.sync()
            st.timer_start()
            for _ in range(args.steps):
                one_step(st, hz, prog_z)
            z_ms = st.timer_stop() / args.steps
            z_norm = st.norm2()
        if abs(z_norm - 1.0) > (1e-9 if dtype == "complex128" else 1e-4):
            raise SystemExit(f"bench: zero-support run norm {z_norm} != 1")
        zs = {"ms_per_step": z_ms, "value": updates_per_step / (z_ms * 1e-3), "unit": UNIT,
              "tiles_visited_fraction_per_pass": [1.0 if s_.desc.n_active < 0 else 2.0 ** (s_.desc.n_active - (n - prog_z.stats["tile_bits"]))
                                                  for s_ in prog_z.passes],
              "what": "same circuit and kernels; passes skip the tiles that are provably still zero because the run starts "
                      "from |0...0> (index bits of qubits no pass has touched yet are 0). Exact; not used for `value`."}
    pass_ms = [ms for ms, kind, _ in per_launch if kind == 10]
    streamed_ms, init_ms = split_init_pass(per_launch, args.steps, prog.fused_init)
    avg_pass_ms = float(np.mean(streamed_ms)) if streamed_ms else float("nan")
    alg_bytes = 2 * amp_bytes * (1 << n)                       # one read + one write of the state
    peak, peak_src = _peaks()
    achieved = alg_bytes / (avg_pass_ms * 1e-3) / 1e9
    pass_share = sum(streamed_ms) / total_ms if total_ms else None
    traffic, traffic_src = None, None
    tf = ROOT / "profiles" / "r02" / "traffic_k_pass_jit_n30.json"
    if tf.exists() and n == 30 and dtype == "complex128" and WORKLOAD == "random":
        t_ = json.loads(tf.read_text())
        traffic, traffic_src = t_["dram_bytes_per_launch"], t_["source"]

    # ---- the other BASELINE workloads on this GPU (short runs, beside the headline) ----
    others = None
    if not args.no_others and WORKLOAD == "random" and n == 30 and dtype == "complex128":
        others = [_other_workload(args, wl, n_, dt_) for wl, n_, dt_ in
                  (("qft", 28, "complex128"), ("ghz", 30, "complex128"), ("ghz", 20, "complex128"), ("random", 30, "complex64"))]

    # ---- the per-gate ABI (a1/a2 operator face) and the observables against the same roofline ----
    per_gate = None
    if not args.no_others and n >= 20:
        try:                                         # in a child process: a secondary table never costs the headline line
            r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--per-gate-child", "--qubits", str(n), "--dtype", dtype,
                                "--device", str(args.device)], capture_output=True, text=True, timeout=240)
            rows = json.loads(r.stdout.strip().splitlines()[-1])
            per_gate = {"what": "quantum_simulations_b200.bench.kernel (quick): one launch per gate, CUDA events, algorithmic "
                                "bytes (2 x state per gate, half for a controlled gate, 1 x for an observable) / time / measured HBM peak",
                        "rows": [{k: r_[k] for k in ("kernel", "what", "ms", "gbs", "frac")} for r_ in rows]}
        except Exception as e:
            per_gate = {"error": f"{type(e).__name__}: {e}"[:200]}

    # ---- end to end through the public API, result in pinned HOST memory ----
    e2e = None
    if not args.no_e2e:
        host = PinnedBuffer((1 << n) * amp_bytes)
        out = host.array(dtype, 1 << n)
        simulate(cd, dtype=dtype, device=args.device, out=out, **ckw)       # warm-up
        reps = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        phases: dict = {}
        for _ in range(reps):
            simulate(cd, dtype=dtype, device=args.device, out=out, phases=phases, **ckw)
        e2e_s = (time.perf_counter() - t0) / reps
        import ctypes
        from quantum_simulations_b200 import _lib as L
        prog_bytes = n_pass * ctypes.sizeof(L.QsvPass) + prog.stats["micro_ops"] * ctypes.sizeof(L.QsvOp)
        e2e = {"value": updates_per_step / e2e_s, "unit": UNIT, "ms_per_step": e2e_s * 1e3,
               "h2d_bytes_per_step": int(prog_bytes), "d2h_bytes_per_step": int((1 << n) * amp_bytes),
               "what": "kernel.cuda_dense.simulate(circuit, out=pinned host array): validate + compile + "
                       "cudaMalloc + |0> + passes + D2H of the full state, host perf_counter"}
        nrm = float(np.vdot(out[: 1 << 20], out[: 1 << 20]).real)
        e2e["host_prefix_norm"] = nrm
        with DeviceState(n, dtype, args.device) as st2:           # what the PCIe copy alone costs on this box
            st2.init_zero(); st2.sync()
            t0 = time.perf_counter()
            st2.download(out)
            d2h_s = time.perf_counter() - t0
        e2e["phases_ms"] = {k: round(v / reps, 2) for k, v in phases.items()}
        e2e["d2h_only_ms"] = d2h_s * 1e3
        e2e["d2h_gbs"] = (1 << n) * amp_bytes / d2h_s / 1e9
        host.free()
        # the same call in a FRESH process with the on-disk kernel cache switched off: a circuit structure
        # never seen before pays NVRTC for each of its passes (in parallel) on top of the warm number
        try:
            r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--e2e-cold-child", "--qubits", str(n), "--dtype", dtype,
                                "--device", str(args.device), "--workload", WORKLOAD],
                               capture_output=True, text=True, timeout=300, env=dict(os.environ, QSV_JIT_CACHE="0"))
            cold = json.loads(r.stdout.strip().splitlines()[-1])
            e2e["cold"] = {"ms_per_step": cold["ms"], "value": updates_per_step / (cold["ms"] * 1e-3), "unit": UNIT,
                           "jit": cold["jit"], "what": "first simulate() of a fresh process with QSV_JIT_CACHE=0 (every pass kernel "
                                                       "compiled by NVRTC), pinned host output, host perf_counter"}
        except Exception as e:      # context only
            e2e["cold"] = {"error": str(e)[:200]}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "config": {**info, "passes_per_step": n_pass, "rounds_per_step": prog.stats["rounds"],
                   "ops_per_step": prog.stats["micro_ops"], "per_pass_ms": [round(x, 3) for x in pass_ms[-n_pass:]],
                   "per_pass_ops": [s_.n_micro_ops for s_ in prog.passes],
                   "per_pass_rounds": [s_.desc.n_rounds for s_ in prog.passes],
                   "tile_bits": prog.stats["tile_bits"], "low_bits": prog.stats["low_bits"],
                   "init": ("fused into the first pass (qsv_pass.zero_input: it reads nothing; the shard is zero-filled and only "
                            "the tile that holds amplitude 0 is computed, every other tile is zero before and after a linear "
                            "pass) — per_pass_ms[0]; not part of the roofline average") if prog.fused_init
                           else "cudaMemset + set amp[0] before the passes",
                   "init_pass_ms": round(float(np.mean(init_ms)), 3) if init_ms else None, "init_note": init_note,
                   "planner_switches": {"low_store_round": not args.no_low_store_round,
                                        "warp_local_rounds": bool(args.warp_local_rounds),
                                        "streaming_stores": os.environ.get("QSV_JIT_STCS") == "1",
                                        "init_pass_full": "QSV_INIT_PASS_FULL" in os.environ,
                                        "tile_block_log2": os.environ.get("QSV_JIT_TILE_BLOCK", "default"),
                                        "pair_loads": os.environ.get("QSV_JIT_PAIR", "default")},
                   "l2_hygiene": f"state {(1 << n) * amp_bytes / 2**30:.0f} GiB >> 126 MB L2: every pass streams from HBM",
                   "host_compile_s": compile_s,
                   "state_fingerprint": fingerprints[0] if fingerprints else None},
        "gate_layers_per_s": info["levels"] / (ms_per_step * 1e-3),
        "hbm_gbs_per_gate_layer": info["levels"] * alg_bytes / (ms_per_step * 1e-3) / 1e9,
        "roofline": {"bound": "hbm", "kernel": "k_pass_jit (run-time specialised ring kernel; k_pass_ring / k_pass interpret "
                                                 "the passes NVRTC could not build)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src, "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": avg_pass_ms,
                     "launches_timed": len(streamed_ms), "share_of_step": pass_share,
                     "launches": "every pass that reads and writes the state once (the write-only init pass is excluded)"},
        "other_workloads": others,
        "per_gate_kernels": per_gate,
        "zero_support_skipping": zs,
        "full_first_pass": full_first,
        "jit": jit_stats(),
        "gpu_launches": len(per_launch) + (0 if prog.fused_init else 2 * args.steps),   # + memset & set-amp of |0> unless fused
        "clocks": clk,
        "e2e": e2e,
    }
    if not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(args.cpu_qubits)
    print(json.dumps(line), flush=True)


BASELINE_QUBITS = {2: 34, 4: 34, 8: 36}     # BASELINE.json configs[3] (34 qubits on 2/4 B200) and configs[4] (36 on 8)


def _timed_sharded(sim, cd, args, steps, warmup, ckw, tol):
    """Plan + prepare + warm-up + `steps` timed executions of `cd` on the sharded simulator.
    Returns a dict with the device-timed step (max over ranks), per-launch timings of rank 0 and the plan."""
    from quantum_simulations_b200.circuit.passes import SwapStep
    dist, st = sim.dist, sim.shard.state
    t0 = time.perf_counter()
    prog = sim.plan(cd, **ckw)
    compile_s = time.perf_counter() - t0
    sim.prepare(prog)

    def global_norm() -> float:
        return dist.allreduce(float(st.norm2()), "sum")

    sim.run(prog)
    plan_note = "default planner options"
    if abs(global_norm() - 1.0) > tol:
        # insurance: a plan that does not even conserve the norm is replaced by the conservative planner
        # (swaps on the top positions after a relabel pass, no eager flips, no table phases) before timing
        sim.shard.release(prog)
        prog = sim.plan(cd, swap_anywhere=False, rank_flips=False, eager_flips=False, table_phases=False, **ckw)
        sim.prepare(prog)
        plan_note = "FALLBACK: conservative planner options (the default plan failed the norm check)"
        sim.run(prog)
    for _ in range(max(warmup - 1, 0)):
        sim.run(prog)
    st.sync(); dist.barrier()
    clocks = ClockSampler(sim.local_rank).start() if sim.rank == 0 else None
    st.timing(True)
    st.timer_start()
    for _ in range(steps):
        sim.run(prog)
    total_ms = st.timer_stop()
    per_launch = st.take_timings()
    st.timing(False)
    dist.barrier()
    clk = clocks.stop() if clocks else None
    total_ms = dist.allreduce(float(total_ms), "max")
    nrm = global_norm()
    if abs(nrm - 1.0) > tol:
        raise SystemExit(f"bench: state norm {nrm} != 1 — result invalid")
    seq = "".join(("S%d" % len(s_.global_bits)) if isinstance(s_, SwapStep) else ("P" if s_.n_micro_ops else "p") for s_ in prog.steps)
    return {"prog": prog, "total_ms": total_ms, "per_launch": per_launch, "clocks": clk, "norm": nrm,
            "compile_s": compile_s, "plan_note": plan_note, "sequence": seq}


def _cross_g_parity(args, world, rank, local_rank, dist, n_par: int = 30, samples: int = 1 << 20) -> dict:
    """OUTSIDE the timed region: the same circuit family at n_par qubits, once sharded over all ranks (the path
    that was timed: stage plan, pipelined swaps) and once on rank 0's GPU alone; 2^20 seeded amplitude indices
    are compared.  (Full-vector parity against the CPU oracle is the test suite's job, n <= 26.)"""
    from quantum_simulations_b200.runner.multi_gpu import ShardedSimulator
    cd, _ = workload(n_par)
    g = world.bit_length() - 1
    sim = ShardedSimulator(n_par, args.dtype)
    shard = sim.simulate(cd)
    logical, pipelined = sim.logical_rank, sim.shard.pipelined_swaps
    sim.close()
    idx = np.random.default_rng(2026).integers(0, 1 << n_par, size=samples, dtype=np.int64)
    n_loc = n_par - g
    mine = (idx >> n_loc) == logical
    box = dist.all_gather_object((np.nonzero(mine)[0], shard[idx[mine] & ((1 << n_loc) - 1)]))
    del shard
    out = None
    if rank == 0:
        from quantum_simulations_b200.kernel.cuda_dense import simulate
        got = np.zeros(samples, dtype=np.complex128)
        seen = np.zeros(samples, dtype=bool)
        for where, vals in box:
            got[where] = vals
            seen[where] = True
        one = simulate(cd, dtype=args.dtype, device=local_rank)
        err = float(np.abs(got - one[idx]).max())
        tol = 1e-12 if args.dtype == "complex128" else 1e-5
        out = {"what": f"{samples} seeded amplitude indices of the {n_par}-qubit circuit of the same family: sharded run on "
                       f"{world} GPUs (same planner and swap path as the timed run) vs the single-GPU run on rank 0's GPU",
               "n_qubits": n_par, "sampled_amplitudes": int(samples), "all_indices_covered": bool(seen.all()),
               "pipelined_swaps_in_the_sharded_run": int(pipelined), "max_abs_diff": err, "tolerance": tol,
               "ok": bool(seen.all() and err <= tol)}
    dist.barrier()
    return out


def _one_gpu_rate(args, local_rank: int, n: int = 30, steps: int = 5, warmup: int = 3) -> dict:
    """The 1-GPU rate of the same circuit family (BASELINE configs[2], n = 30) measured in THIS run on rank 0's GPU:
    the denominator of SURVEY.md section 8d's parallel efficiency."""
    from quantum_simulations_b200.circuit.sharding import plan_single
    from quantum_simulations_b200.kernel.cuda import DeviceState
    from quantum_simulations_b200.kernel.cuda_dense import circuit_ops
    cd, info = workload(n)
    prog = plan_single(circuit_ops(cd), n, args.dtype, True, False)
    with DeviceState(n, args.dtype, local_rank) as st:
        h = st.upload_program(prog)
        for _ in range(warmup):
            st.replay(h)
        st.sync()
        st.timer_start()
        for _ in range(steps):
            st.replay(h)
        ms = st.timer_stop() / steps
    return {"n_qubits": n, "ms_per_step": ms, "value": info["gates"] * (1 << n) / (ms * 1e-3), "unit": UNIT,
            "steps": steps, "warmup": warmup}


def bench_multi(args) -> None:
    """N > 1: one process per GPU (launched by torchrun; the processes talk over runner/plumbing.py and NVLink).  Default workload = BASELINE.json configs[3] / configs[4]: the random
    depth-20 circuit at 34 qubits on 2 and 4 GPUs and at 36 qubits (1 TiB) on 8, sharded by the top log2(N) qubits
    (--qubits overrides).  Beside the headline the line carries the WEAK series (2^30 amplitudes per GPU,
    n = 30 + log2 N), a cross-G parity check and the 1-GPU rate of this run for the parallel efficiency."""
    import math
    from quantum_simulations_b200 import _lib as L
    from quantum_simulations_b200.runner.multi_gpu import ShardedSimulator

    dtype = args.dtype
    amp_bytes = np.dtype(dtype).itemsize
    world = int(os.environ["WORLD_SIZE"])
    g = int(math.log2(world))
    n = args.qubits if args.qubits is not None else BASELINE_QUBITS.get(world, 30 + g)
    cd, info = workload(n)
    ckw = dict(tile_bits=args.tile_bits, low_bits=args.low_bits, max_rounds=args.max_rounds)
    if args.no_low_store_round:
        ckw["low_store_round"] = False
    tol = 1e-9 if dtype == "complex128" else 1e-4

    # ---- parity FIRST: the sharded path (stage plan, pipelined swaps, the data-movement switches of the pass kernels)
    # against the single-GPU run, before anything is timed.  If it fails under the default switches, the paired loads /
    # tile blocks of the pass kernels are switched off — the setting every multi-GPU line of profiles/r02 was measured
    # with — checked again, and the run is timed under that; the line says so (config.preflight).
    parity, preflight = None, None
    if not args.no_parity:
        from quantum_simulations_b200.runner.multi_gpu import dist_env
        from quantum_simulations_b200.runner.plumbing import init_plumbing
        rank_, local_rank_, _ = dist_env()
        dist_ = init_plumbing()
        parity = _cross_g_parity(args, world, rank_, local_rank_, dist_, min(30, n))
        # what this path runs by default without having met the hardware yet, most recent first; each level is
        # switched off (on every rank) only if the check before it failed and the user has not set it explicitly
        fallbacks = (({"QSV_JIT_PAIR": "0", "QSV_JIT_TILE_BLOCK": "0"}, "paired loads / tile blocks of the pass kernels off"),
                     ({"QSV_PLAN_SEARCH": "0"}, "greedy stage plan (no seeded search)"))
        failed = []
        for env_, what_ in fallbacks:
            ok = dist_.broadcast_object(None if parity is None else bool(parity["ok"]), src=0)
            if ok or any(k in os.environ for k in env_):
                break
            os.environ.update(env_)
            L.load().qsv_release_cached()
            failed.append({"switched_off_after": what_, "env": env_, "failed_check": parity})
            parity = _cross_g_parity(args, world, rank_, local_rank_, dist_, min(30, n))
        if failed:
            preflight = {"what": "FALLBACK: the parity check failed under the default switches; the run is timed with "
                                 + " ".join(f"{k}={v}" for f_ in failed for k, v in f_["env"].items()),
                         "first_check": failed[0]["failed_check"], "levels": failed}
        L.load().qsv_release_cached()
    sim = ShardedSimulator(n, dtype, fused_exchange=True) if args.fused_exchange else ShardedSimulator(n, dtype)
    rank, dist = sim.rank, sim.dist
    updates_per_step = len(cd["gates"]) * (1 << n)
    run = _timed_sharded(sim, cd, args, args.steps, args.warmup, ckw, tol)
    prog, total_ms, per_launch, clk = run["prog"], run["total_ms"], run["per_launch"], run["clocks"]

    # end to end through the public object: plan + (cached) specialisation + run + D2H of the WHOLE shard, streamed
    # through two pinned staging buffers (the state does not fit in host memory at 34 / 36 qubits)
    n_loc = n - g
    shard_bytes = amp_bytes * (1 << n_loc)
    e2e_s, e2e_sum = None, 0.0
    if not args.no_e2e:
        acc = [0.0]

        def sink(view, first):          # touch every piece on the host: first amplitude of each (a cheap witness)
            acc[0] += float(abs(view[0]))

        sim.simulate(cd, sink=sink, **ckw)
        dist.barrier()
        reps = max(1, min(args.steps, 2 if n_loc > 31 else 3))
        t0 = time.perf_counter()
        for _ in range(reps):
            sim.simulate(cd, sink=sink, **ckw)
        dist.barrier()
        e2e_s, e2e_sum = dist.allreduce((time.perf_counter() - t0) / reps, "max"), acc[0]
    peer_swap, peer_error, xchg_sms, pipeline_on = sim.peer_swap, sim.shard.peer_error, getattr(sim.shard, "xchg_sms", None), bool(getattr(sim.shard, "pipeline", False))
    fused_flag = bool(getattr(sim, "fused_exchange", False))
    sim.close()
    L.load().qsv_release_cached()
    dist.barrier()

    # ---- beside the headline (all outside its timed region) ----
    weak = None
    if not args.no_weak and n != 30 + g:
        wsim = ShardedSimulator(30 + g, dtype)
        wcd, winfo = workload(30 + g)
        w = _timed_sharded(wsim, wcd, args, min(args.steps, 5), min(args.warmup, 3), ckw, tol)
        wsim.close()
        L.load().qsv_release_cached()
        w_ms = w["total_ms"] / min(args.steps, 5)
        weak = {"n_qubits": 30 + g, "amps_per_gpu_log2": 30, "gates": winfo["gates"], "ms_per_step": w_ms,
                "value": winfo["gates"] * (1 << (30 + g)) / (w_ms * 1e-3), "unit": UNIT, "step_sequence": w["sequence"],
                "what": "weak-scaling series: 2^30 amplitudes per GPU, same circuit family"}
    one = _one_gpu_rate(args, sim.local_rank) if (rank == 0 and not args.no_parity) else None

    if rank == 0:
        ms_per_step = total_ms / args.steps
        value = updates_per_step / (ms_per_step * 1e-3)
        pass_ms, init_ms = split_init_pass([x for x in per_launch if x[1] != 11], args.steps, prog.fused_init)
        swap_ms = [(ms, kind - 20) for ms, kind, _ in per_launch if 20 <= kind < 30]
        scatter_ms = [(ms, kind - 40) for ms, kind, _ in per_launch if 40 <= kind < 50]
        piped = [(ms, kind - 50, pi) for ms, kind, pi in per_launch if 50 <= kind < 60]
        avg_pass_ms = float(np.mean(pass_ms)) if pass_ms else float("nan")
        alg_bytes = 2 * amp_bytes * (1 << n_loc)
        peak, peak_src = _peaks()
        achieved = alg_bytes / (avg_pass_ms * 1e-3) / 1e9
        per_step = lambda xs: xs[-max(1, len(xs) // max(args.steps, 1)):] if xs else []      # noqa: E731
        nv = [{"bits": b, "ms": round(ms, 3),
               "sent_gb_per_gpu": round((1 - 0.5 ** b) * shard_bytes / 1e9, 3),
               "gbs_per_direction": round((1 - 0.5 ** b) * shard_bytes / (ms * 1e-3) / 1e9, 1),
               "frac_of_900": round((1 - 0.5 ** b) * shard_bytes / (ms * 1e-3) / 900e9, 3)} for ms, b in per_step(swap_ms)]
        # a pipelined region = its passes + the exchange beside them; what the exchange ADDS to the step is the
        # region minus what its passes take on their own (the average streaming pass of this run)
        regions = []
        for ms, b, pi in per_step(piped):
            n_p = pi // 16 + pi % 16
            sent = (1 - 0.5 ** b) * shard_bytes
            regions.append({"bits": b, "ms": round(ms, 3), "passes_before": pi // 16, "passes_after": pi % 16,
                            "sent_gb_per_gpu": round(sent / 1e9, 3),
                            "exchange_alone_at_900_ms": round(sent / 900e9 * 1e3, 3),
                            "exposed_ms": round(ms - n_p * avg_pass_ms, 3),
                            "nvlink_gbs_per_direction_over_region": round(sent / (ms * 1e-3) / 1e9, 1)})
        exposed = sum(r_["exposed_ms"] for r_ in regions) + sum(ms for ms, _ in per_step(swap_ms)) + 0.0
        import ctypes
        prog_bytes = len(prog.passes) * ctypes.sizeof(L.QsvPass) + prog.stats["micro_ops"] * ctypes.sizeof(L.QsvOp)
        eff = None
        if one is not None:
            eff = {"definition": "SURVEY.md 8d: amp-updates/s at G GPUs / (G x amp-updates/s of the 1-GPU n=30 run of the same "
                                 "circuit family), the 1-GPU rate measured in this run on rank 0's GPU",
                   "value": value / (world * one["value"]), "one_gpu": one,
                   "weak_series_value": None if weak is None else weak["value"] / (world * one["value"])}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": {**info, "amps_per_gpu_log2": n_loc, "sharding": f"top {g} qubits = rank bits",
                       "baseline_config": "BASELINE.json configs[3]/[4]" if n == BASELINE_QUBITS.get(world) else "--qubits override",
                       "passes_per_step": len(prog.passes), "swaps_per_step": prog.stats["swaps"],
                       "swap_bits_per_step": prog.stats["swap_bits"], "ops_per_step": prog.stats["micro_ops"],
                       "step_sequence": run["sequence"],
                       "l2_hygiene": f"shard {(1 << n_loc) * amp_bytes / 2**30:.0f} GiB >> 126 MB L2",
                       "host_compile_s": run["compile_s"], "plan": run["plan_note"],
                       "plan_search": prog.stats.get("search"), "preflight": preflight,
                       "scaling_note": "amplitudes per GPU: 2^30 at N=1 (configs[2]), 2^33 / 2^32 / 2^33 at N=2 / 4 / 8 "
                                       "(configs[3], [4]); the weak series with 2^30 per GPU is under weak_series",
                       "timing": "CUDA events on each rank's stream, max over ranks"},
            "gate_layers_per_s": info["levels"] / (ms_per_step * 1e-3),
            "hbm_gbs_per_gate_layer_per_gpu": info["levels"] * alg_bytes / (ms_per_step * 1e-3) / 1e9,
            "roofline": {"bound": "hbm", "kernel": "k_pass_jit / k_pass_ring", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src, "traffic": None,
                         "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": avg_pass_ms,
                         "launches_timed": len(pass_ms), "share_of_step": sum(pass_ms) / total_ms,
                         "launches": "every whole-shard pass launch that reads and writes the shard once (the write-only init "
                                     f"pass, {round(float(np.mean(init_ms)), 3) if init_ms else None} ms, and the chunked "
                                     "launches inside pipelined regions are excluded)"},
            "pipelined_swap": {"enabled": pipeline_on, "count_per_step": len(piped) // max(args.steps, 1), "regions": regions,
                               "xchg_sms": xchg_sms,
                               "what": "stage transition executed chunk by chunk (qsv_swap_pipelined): the TMA exchange kernel of "
                                       "chunk j on xchg_sms SMs beside the pass kernels of the neighbouring chunks on the other SMs; "
                                       "ms = the whole region (its passes included)"},
            "scatter_pass": {"enabled": fused_flag, "count_per_step": len(scatter_ms) // max(args.steps, 1),
                             "ms": [round(ms, 3) for ms, _ in per_step(scatter_ms)]},
            "nvlink": {"path": "peer memory over CUDA IPC: TMA bulk-copy exchange kernel, flag-word ordering (csrc/xchg.cuh)" if peer_swap
                       else f"chunked ncclSend/ncclRecv ({peer_error or 'QSV_SWAP=nccl'})", "swaps": nv,
                       "exposed_ms_per_step": round(exposed, 3), "share_of_step": exposed / ms_per_step,
                       "pipelined_regions_share_of_step": sum(ms for ms, _, _ in per_step(piped)) / ms_per_step,
                       "peak_gbs_per_direction": 900.0},
            "parallel_efficiency": eff, "weak_series": weak, "parity": parity,
            "gpu_launches": len(per_launch) + (0 if prog.fused_init else 2 * args.steps),
            "clocks": clk,
            "e2e": None if e2e_s is None else
                   {"value": updates_per_step / e2e_s, "unit": UNIT, "ms_per_step": e2e_s * 1e3,
                    "h2d_bytes_per_step": int(prog_bytes) * world, "d2h_bytes_per_step": int(shard_bytes) * world,
                    "d2h_gbs_per_gpu": shard_bytes / max(e2e_s - ms_per_step * 1e-3, 1e-9) / 1e9, "host_witness": e2e_sum,
                    "what": "ShardedSimulator.simulate(circuit, sink=...) on every rank: validate + stage planning + (cached) "
                            "kernel specialisation + |0> + passes + swaps + D2H of the WHOLE shard, streamed through two "
                            "pinned staging buffers of 256 MiB to a host sink (the state is larger than host memory at "
                            "34 / 36 qubits); host perf_counter, max over ranks"},
        }
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.close()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--qubits", type=int, default=None)
    ap.add_argument("--workload", default="random", choices=["random", "qft", "ghz"])
    ap.add_argument("--dtype", default="complex128", choices=["complex64", "complex128"])
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--tile-bits", type=int, default=None)
    ap.add_argument("--low-bits", type=int, default=None)
    ap.add_argument("--max-rounds", type=int, default=3)
    ap.add_argument("--defer-diagonals", action="store_true")
    ap.add_argument("--no-fold-tables", action="store_true")
    ap.add_argument("--cpu-qubits", type=int, default=20)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-zero-support", action="store_true")
    ap.add_argument("--e2e-cold-child", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--per-gate-child", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-others", action="store_true", help="N = 1: skip the short runs of the other BASELINE workloads")
    ap.add_argument("--no-low-store-round", action="store_true",
                    help="experiment: no idle round before stores whose registers hold a low (row) position")
    ap.add_argument("--low-store-bits", type=int, default=None,
                    help="experiment: idle round before the stores only if a store position below this is in registers")
    ap.add_argument("--warp-local-rounds", action="store_true",
                    help="experiment (N = 1): warp-local round exchanges with __syncwarp() instead of the group barrier")
    ap.add_argument("--streaming-stores", action="store_true",
                    help="experiment (N = 1): st.global.cs (evict-first) for the final stores of a pass")
    ap.add_argument("--no-weak", action="store_true", help="N > 1: skip the weak-series run beside the headline")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the cross-G parity check and the 1-GPU rate")
    ap.add_argument("--fused-exchange", action="store_true",
                    help="N > 1: second buffer per shard, the pass before a swap stores straight into the peers (qsv_pass_scatter)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    global WORKLOAD
    WORKLOAD = args.workload
    if args.impl == "reference":
        reference_arm(args)
        return
    if args.per_gate_child:
        import io as _io
        from quantum_simulations_b200.bench.kernel import bench_kernel
        rows_ = bench_kernel(args.qubits or 30, args.dtype, reps=3, device=args.device, out=_io.StringIO(), quick=True)
        print(json.dumps(rows_), flush=True)
        return
    if args.e2e_cold_child:
        from quantum_simulations_b200.kernel.cuda_dense import simulate
        from quantum_simulations_b200.storage.pinned import PinnedBuffer
        n_ = args.qubits or 30
        cd_, _ = workload(n_)
        host_ = PinnedBuffer((1 << n_) * np.dtype(args.dtype).itemsize)
        out_ = host_.array(args.dtype, 1 << n_)
        t0_ = time.perf_counter()
        simulate(cd_, dtype=args.dtype, device=args.device, out=out_)
        print(json.dumps({"ms": (time.perf_counter() - t0_) * 1e3, "jit": jit_stats()}), flush=True)
        return
    world = int(os.environ.get("WORLD_SIZE", args.gpus))
    if world == 1:
        if args.qubits is None:
            args.qubits = 30
        bench_single(args)
    else:
        bench_multi(args)


if __name__ == "__main__":
    main()
