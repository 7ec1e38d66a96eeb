#!/bin/bash
# BASELINE.json configs[4]: random depth-20, 36 qubits complex128 (1 TiB state) on 8 x B200
cd "$(dirname "$0")/.."
out=gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus 8 --qubits 36 --steps 2 --warmup 1 --no-e2e > $out/bench_n36_8gpu.log 2>$out/bench_n36_8gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
f='gpurun_out/bench_n36_8gpu.log'
try:
    r=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, 'n',r['config']['n_qubits'], round(r['ms_per_step'],2), r['config']['step_sequence'], r['nvlink'], 'pass avg', round(r['roofline']['avg_launch_ms'],2), 'frac', round(r['roofline']['frac'],3))
except Exception as e:
    print(f, 'ERR', e, open(f).read()[-500:], open(f.replace('.log','.err')).read()[-2500:])
PY
nvidia-smi --query-gpu=memory.used,memory.total --format=csv | head -3
