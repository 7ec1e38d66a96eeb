#!/bin/bash
cd "$(dirname "$0")/.."
out=gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $out/gpu.txt
timeout 600 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > $out/bench_full.log 2>$out/bench_full.err; echo "bench rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name k_pass_ring --launch-skip 9 --launch-count 1 -o $out/prof_r01c -f python tools/ncu_cases.py 28 > $out/ncu_r01c.log 2>&1; echo "ncu rc=$?"
tail -c 600 $out/bench_full.log
