#!/bin/bash
# N-GPU call: parity across real shards (world up to N), then the weak-scaling bench line at N
N=${1:-4}
cd "$(dirname "$0")/.."
out=gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -s > $out/pytest_mgpu_$N.log 2>&1; echo "pytest rc=$?"; grep -E "max\|d\||passed|failed|chunk files" $out/pytest_mgpu_$N.log | tail -24
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 3 --warmup 3 > $out/bench_n$N.log 2>$out/bench_n$N.err; echo "bench rc=$?"
python - $N <<'PY'
import json,sys
f='gpurun_out/bench_n%s.log'%sys.argv[1]
try:
    r=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, 'n',r['config']['n_qubits'], round(r['ms_per_step'],2), r['config']['step_sequence'], r['nvlink'], r.get('overlapped_pass_swap'), 'pass avg', round(r['roofline']['avg_launch_ms'],2), 'e2e ms', r['e2e'] and round(r['e2e']['ms_per_step'],1))
except Exception as e:
    print(f, 'ERR', e, open(f).read()[-500:], open(f.replace('.log','.err')).read()[-2500:])
PY
