#!/bin/bash
# BASELINE.json configs[3]: random depth-20, 34 qubits complex128 on N = 2 or 4 B200
N=${1:-4}
cd "$(dirname "$0")/.."
out=gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus $N --qubits 34 --steps 2 --warmup 1 --no-e2e > $out/bench_n34_${N}gpu.log 2>$out/bench_n34_${N}gpu.err; echo "bench rc=$?"
python - $N <<'PY'
import json,sys
f='gpurun_out/bench_n34_%sgpu.log'%sys.argv[1]
try:
    r=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, 'n',r['config']['n_qubits'], round(r['ms_per_step'],2), r['config']['step_sequence'], r['nvlink'], 'pass avg', round(r['roofline']['avg_launch_ms'],2), 'frac', round(r['roofline']['frac'],3))
except Exception as e:
    print(f, 'ERR', e, open(f).read()[-500:], open(f.replace('.log','.err')).read()[-2500:])
PY
