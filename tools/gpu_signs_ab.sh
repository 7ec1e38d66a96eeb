#!/bin/bash
# A/B of the round-2 generator changes (deferred sign flips, additive shared-memory offsets) and of the
# store-round threshold of the planner; parity of the new kernels first.
cd "$(dirname "$0")/.."
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $out/pytest_parity_signs.log 2>&1; echo "parity rc=$?"; tail -2 $out/pytest_parity_signs.log
run() {  # tag, env, flags
  env $2 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-zero-support --no-others $3 > $out/bench_ab_$1.log 2>$out/bench_ab_$1.err; echo "bench $1 rc=$?"
  python - $out/bench_ab_$1.log <<'PY'
import json, sys
try:
    r = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); c = r["config"]
    print(sys.argv[1], "ms/step", round(r["ms_per_step"], 2), "per pass", c["per_pass_ms"], "rounds", c["per_pass_rounds"], "roofline", round(r["roofline"]["frac"], 3))
except Exception as e:
    print(sys.argv[1], "ERR", e, open(sys.argv[1].replace(".log", ".err")).read()[-800:])
PY
}
run deferred "QSV_X=0" ""
run executed "QSV_JIT_SIGNS=0" ""
run deferred_lsb2 "QSV_X=0" "--low-store-bits 2"
run deferred_lsb1 "QSV_X=0" "--low-store-bits 1"
run deferred_nbuf7 "QSV_JIT_NBUF=7" ""
run deferred_c64 "QSV_X=0" "--dtype complex64"
run executed_c64 "QSV_JIT_SIGNS=0" "--dtype complex64"
