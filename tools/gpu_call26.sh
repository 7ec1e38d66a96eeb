#!/bin/bash
cd "$(dirname "$0")/.."
out=gpurun_out
for cfg in "4 4" "8 4" "2 4" "4 8" "8 8" "4 2" "16 2"; do
  set -- $cfg
  QSV_SWAP_GRID=$1 QSV_SWAP_UNROLL=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 3 --warmup 2 --no-e2e > $out/bench_swap_$1_$2.log 2>$out/bench_swap_$1_$2.err
  python - $1 $2 <<'PY'
import json,sys
f='gpurun_out/bench_swap_%s_%s.log'%(sys.argv[1],sys.argv[2])
try:
    r=json.loads(open(f).read().strip().splitlines()[-1])
    print(sys.argv[1:], round(r['ms_per_step'],2), r['nvlink']['swaps'])
except Exception as e:
    print(f,'ERR',e, open(f.replace('.log','.err')).read()[-800:])
PY
done
