#!/bin/bash
# 2-GPU call: parity across real shards, then the weak-scaling bench line at N=2
cd "$(dirname "$0")/.."
out=gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > $out/gpus2.txt
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > $out/pytest_mgpu.log 2>&1; echo "pytest rc=$?"; tail -15 $out/pytest_mgpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 3 --warmup 3 > $out/bench_n2.log 2>$out/bench_n2.err; echo "bench2 rc=$?"
tail -c 1500 $out/bench_n2.log; tail -5 $out/bench_n2.err
