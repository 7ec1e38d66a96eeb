#!/bin/bash
cd "$(dirname "$0")/.."
out=gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 > $out/bench_r01_final.log 2>$out/bench_r01_final.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --max-rounds 3 --no-cpu --no-e2e > $out/bench_mr3.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --max-rounds 4 --no-cpu --no-e2e > $out/bench_mr4.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --max-rounds 5 --no-cpu --no-e2e > $out/bench_mr5.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_r01.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > $out/ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name k_pass_jit --launch-skip 8 --launch-count 8 -o $out/prof_r01_jit -f python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > $out/ncu_jit.log 2>&1; echo "ncu full rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_mr*.log'))+['gpurun_out/bench_r01_final.log']:
    try:
        r=json.loads(open(f).read().strip().splitlines()[-1]); c=r['config']
        print(f.split('/')[-1], round(r['ms_per_step'],2),'ms/step passes',c['passes_per_step'],'rounds',c['rounds_per_step'],'ops',c.get('ops_per_step'),'frac',round(r['roofline']['frac'],3), 'e2e', r.get('e2e') and round(r['e2e']['ms_per_step'],1))
        print('   ms',c.get('per_pass_ms'))
    except Exception as e:
        print(f,'ERR',e, open(f).read()[-300:])
PY
