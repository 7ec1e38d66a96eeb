#!/bin/bash
# the other single-GPU BASELINE configs: QFT-28 (configs[1]), QFT-30, GHZ-30, complex64
cd "$(dirname "$0")/.."
out=gpurun_out
run() { name=$1; shift; timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e "$@" > $out/bench_$name.log 2>$out/bench_$name.err; }
run qft28 --workload qft --qubits 28
run qft30 --workload qft --qubits 30
run ghz30 --workload ghz --qubits 30
run rnd30_c64 --dtype complex64
run qft28_c64 --workload qft --qubits 28 --dtype complex64
python - <<'PY'
import json
for f in ['qft28','qft30','ghz30','rnd30_c64','qft28_c64']:
    f='gpurun_out/bench_%s.log'%f
    try:
        r=json.loads(open(f).read().strip().splitlines()[-1]); c=r['config']
        print(f.split('/')[-1], round(r['ms_per_step'],2),'ms/step passes',c['passes_per_step'],'rounds',c['rounds_per_step'],'ops',c.get('ops_per_step'),'frac',round(r['roofline']['frac'],3), 'gates', c['gates'], 'levels', c['levels'], 'layerGB/s', round(r['hbm_gbs_per_gate_layer']))
        print('   ms',c.get('per_pass_ms'))
    except Exception as e:
        print(f,'ERR',e, open(f).read()[-300:], open(f.replace('.log','.err')).read()[-800:])
PY
