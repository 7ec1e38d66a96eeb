#!/bin/bash
cd "$(dirname "$0")/.."
out=gpurun_out
run() { name=$1; shift; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e "$@" > $out/bench_$name.log 2>$out/bench_$name.err; }
QSV_JIT_GROUPS=3 run g3 --max-rounds 3
QSV_JIT_GROUPS=4 run g4 --max-rounds 3
QSV_JIT_GROUPS=4 run g4_c64 --max-rounds 3 --dtype complex64
QSV_JIT_GROUPS=3 run g3_c64 --max-rounds 3 --dtype complex64
QSV_JIT_GROUPS=4 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "specialised or golden" > $out/pytest_g4.log 2>&1; tail -2 $out/pytest_g4.log
python - <<'PY'
import json,glob
for f in ['g3','g4','g3_c64','g4_c64']:
    f='gpurun_out/bench_%s.log'%f
    try:
        r=json.loads(open(f).read().strip().splitlines()[-1]); c=r['config']
        print(f.split('/')[-1], round(r['ms_per_step'],2),'ms/step passes',c['passes_per_step'],'rounds',c['rounds_per_step'],'ops',c.get('ops_per_step'),'frac',round(r['roofline']['frac'],3))
        print('   ms',c.get('per_pass_ms'))
    except Exception as e:
        print(f,'ERR',e, open(f).read()[-300:], open(f.replace('.log','.err')).read()[-800:])
PY
