#!/bin/bash
cd "$(dirname "$0")/.."
out=gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > $out/pytest_mgpu.log 2>&1; echo "pytest rc=$?"; tail -15 $out/pytest_mgpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 3 --warmup 3 > $out/bench_n2_peer.log 2>$out/bench_n2_peer.err; echo "bench2 rc=$?"
QSV_SWAP=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus 2 --steps 3 --warmup 3 > $out/bench_n2_nccl.log 2>$out/bench_n2_nccl.err; echo "bench2 nccl rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/bench_n2_peer.log','gpurun_out/bench_n2_nccl.log'):
    try:
        r=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(r['ms_per_step'],2), r['nvlink'], 'e2e ms', round(r['e2e']['ms_per_step'],1))
    except Exception as e:
        print(f, 'ERR', e, open(f).read()[-500:], open(f.replace('.log','.err')).read()[-1500:])
PY
