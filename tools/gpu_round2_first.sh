#!/bin/bash
# First GPU call of the next round (1 GPU, ~6 min of box time): everything that was changed on the CPU
# after the last GPU minute of round 1 gets validated and timed, with A/B lines for each change.
#   1. pytest -m gpu incl. the opt-in scatter-pass tests on one device
#   2. bench.py default            (first pass = zero-fill + one live tile)
#   3. bench.py QSV_INIT_PASS_FULL (first pass = full zero-input launch, the round-1 measured form)
#   4. bench.py --no-low-store-round (no idle round before low-position stores: 20 rounds instead of 24)
#   5. ncu launch list of (2)
cd "$(dirname "$0")/.."
out=gpurun_out; mkdir -p $out
QSV_TEST_SCATTER=1 timeout 1500 python -m pytest tests -m gpu -q > $out/pytest_gpu_r2.log 2>&1; echo "pytest rc=$?"; grep -E "FAILED|ERROR|passed|failed" $out/pytest_gpu_r2.log | tail -n 12
run() {  # name, env, flags
  env $2 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e $3 > $out/bench_r2_$1.log 2>$out/bench_r2_$1.err; echo "bench $1 rc=$?"
  python - $out/bench_r2_$1.log <<'PY'
import json, sys
try:
    r = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    c = r["config"]
    print(sys.argv[1], "ms/step", round(r["ms_per_step"], 2), "per pass", c["per_pass_ms"], "rounds", c["per_pass_rounds"],
          "init", c.get("init_pass_ms"), c.get("init_note"), "roofline", round(r["roofline"]["frac"], 3))
except Exception as e:
    print(sys.argv[1], "ERR", e, open(sys.argv[1].replace(".log", ".err")).read()[-1500:])
PY
}
run default "QSV_X=0" ""
run init_full "QSV_INIT_PASS_FULL=1" ""
run no_low_store_round "QSV_X=0" "--no-low-store-round"
run warp_local_rounds "QSV_X=0" "--warp-local-rounds"
run both "QSV_X=0" "--no-low-store-round --warp-local-rounds"
run streaming_stores "QSV_X=0" "--streaming-stores"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_r02.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-zero-support > $out/ncu_launches_r02.log 2>&1; echo "ncu list rc=$?"
# 6. do DFMA and DMMA run on separate pipes?  (mixed_mode* sum_tflops above either single peak = yes)
mkdir -p tools/_build && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/fp64_peak tools/fp64_peak.cu \
  && tools/_build/fp64_peak | tee $out/fp64_peak_mixed.json
