#!/bin/bash
cd "$(dirname "$0")/.."
out=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_gpu.log
run() { name=$1; shift; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e "$@" > $out/bench_$name.log 2>$out/bench_$name.err; }
run codes --max-rounds 3
run codes_mr4 --max-rounds 4
python - <<'PY'
import json,glob
for f in ['codes','codes_mr4']:
    f='gpurun_out/bench_%s.log'%f
    try:
        r=json.loads(open(f).read().strip().splitlines()[-1]); c=r['config']
        print(f.split('/')[-1], round(r['ms_per_step'],2),'ms/step passes',c['passes_per_step'],'rounds',c['rounds_per_step'],'ops',c.get('ops_per_step'),'frac',round(r['roofline']['frac'],3))
        print('   ms',c.get('per_pass_ms'))
    except Exception as e:
        print(f,'ERR',e, open(f).read()[-300:], open(f.replace('.log','.err')).read()[-800:])
PY
