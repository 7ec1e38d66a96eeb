#!/bin/bash
cd "$(dirname "$0")/.."
out=gpurun_out
QSV_LIB_NAME=libqsv_g4.so timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tile_shapes or depth20" > $out/pytest_g4.log 2>&1; echo "pytest g4 rc=$?"; tail -2 $out/pytest_g4.log
QSV_LIB_NAME=libqsv_g4.so timeout 90 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --low-bits 3 > $out/bench_g4_a3.log 2>$out/bench_g4_a3.err; echo "bench rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_g4*.log')):
    try:
        r=json.loads(open(f).read().strip().splitlines()[-1]); c=r['config']
        print(f.split('/')[-1], round(r['ms_per_step'],2),'ms/step passes',c['passes_per_step'],'rounds',c['rounds_per_step'],'ops',c.get('ops_per_step'),'frac',round(r['roofline']['frac'],3))
        print('   ms',c.get('per_pass_ms'))
    except Exception as e:
        print(f,'ERR',e, open(f).read()[-300:])
PY
