#!/bin/bash
# Round 2, the very last GPU seconds: the NEW default (paired loads, blocks of 2 tiles) through a wider parity subset,
# the default bench line, and complex64 with / without the switch.
mkdir -p gpurun_out/final
timeout 22 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "specialised_and_interpreted or zero_support or fused_initialisation" > gpurun_out/final/pytest_new_default.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/final/pytest_new_default.log
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-zero-support --no-others"
timeout 12 $B > gpurun_out/final/bench_default.json 2> gpurun_out/final/bench_default.err; echo "bench rc=$?"
timeout 10 $B --dtype complex64 > gpurun_out/final/bench_c64_default.json 2>/dev/null; echo "c64 rc=$?"
QSV_JIT_PAIR=0 QSV_JIT_TILE_BLOCK=0 timeout 10 $B --dtype complex64 > gpurun_out/final/bench_c64_pair0.json 2>/dev/null; echo "c64 pair0 rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/final/bench_*.json")):
    try:
        r = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], round(r["ms_per_step"], 2), round(r["roofline"]["frac"], 3), r["config"]["per_pass_ms"], r["config"]["state_fingerprint"])
    except Exception as e:
        print(f, "no line", e)
PY
