#!/bin/bash
cd "$(dirname "$0")/.."
out=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_gpu.log
for a in 3 4 5; do
  timeout 300 python bench.py --steps 5 --warmup 3 --low-bits $a --no-cpu --no-e2e > $out/bench_a$a.log 2>$out/bench_a$a.err; echo "bench a=$a rc=$?"
done
timeout 300 python bench.py --steps 5 --warmup 3 --low-bits 3 --max-rounds 3 --no-cpu --no-e2e > $out/bench_a3_mr3.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --low-bits 3 --max-rounds 4 --no-cpu --no-e2e > $out/bench_a3_mr4.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_a*.log')):
    try:
        r=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(r['ms_per_step'],2),'ms/step', 'passes',r['config']['passes_per_step'],'rounds',r['config']['rounds_per_step'],'avg pass ms',round(r['roofline']['avg_launch_ms'],2),'frac',round(r['roofline']['frac'],3), 'clk', r['clocks'].get('sm_mhz'))
    except Exception as e:
        print(f,'ERR',e, open(f).read()[-300:])
PY
