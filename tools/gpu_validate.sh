#!/bin/bash
cd "$(dirname "$0")/.."
out=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > $out/bench_full2.log 2>$out/bench_full2.err; echo "bench rc=$?"
python - <<'PY'
import json
r=json.loads(open('gpurun_out/bench_full2.log').read().strip().splitlines()[-1])
print(round(r['ms_per_step'],2), 'e2e', r['e2e']['ms_per_step'], 'roofline', r['roofline']['frac'], r['roofline']['traffic'], 'jit', r['jit'], 'cpu', r['cpu_baseline']['value'])
PY
