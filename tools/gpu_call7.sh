#!/bin/bash
cd "$(dirname "$0")/.."
out=gpurun_out
timeout 400 python tools/sweep_pass.py 30 complex128 > $out/sweep.log 2>&1; echo "sweep rc=$?"
for a in 3 4; do
  timeout 300 python bench.py --steps 5 --warmup 3 --low-bits $a --no-cpu --no-e2e > $out/bench_a$a.log 2>$out/bench_a$a.err; echo "bench a=$a rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_a[34].log')):
    try:
        r=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(r['ms_per_step'],2),'ms/step', 'passes',r['config']['passes_per_step'],'rounds',r['config']['rounds_per_step'],'avg pass ms',round(r['roofline']['avg_launch_ms'],2),'frac',round(r['roofline']['frac'],3), 'clk', r['clocks'].get('sm_mhz'))
    except Exception as e:
        print(f,'ERR',e, open(f).read()[-300:])
PY
timeout 120 python tools/ncu_cases.py 28 > $out/ncu_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass -s 5 -c 5 -f -o $out/prof_ring2 python tools/ncu_cases.py 28 > $out/ncu_run.log 2>&1
echo "ncu rc=$?"; tail -2 $out/ncu_run.log
