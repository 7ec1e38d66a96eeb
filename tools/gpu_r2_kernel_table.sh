#!/bin/bash
# Round 2, last GPU minutes: per-gate kernel / observable roofline table, the new bench-package GPU tests, and the
# tensor-map TMA tile microbenchmark.   gpurun --timeout 140 -- bash tools/gpu_r2_kernel_table.sh
mkdir -p gpurun_out
timeout 30 python -m quantum_simulations_b200.bench.kernel 30 --reps 3 --json gpurun_out/kernel_table_n30.jsonl > gpurun_out/kernel_table_n30_c128.txt 2>&1
echo "kernel table c128 rc=$?"; tail -4 gpurun_out/kernel_table_n30_c128.txt
timeout 45 python -m pytest tests/test_bench_package.py -m gpu -q -x > gpurun_out/pytest_bench_package.log 2>&1
echo "pytest bench package rc=$?"; tail -3 gpurun_out/pytest_bench_package.log
timeout 25 tools/_build/tma_tensor 30 3 6 3,4,5,6,7,8,9,10 22,23,24,25,26,27,28,29 11,12,13,14,15,16,17,18 \
    11,19,20,21,22,23,27,29 9,10,11,12,20,24,25,26 3,5,6,8,10,15,17,18 4,7,13,16,18,24,26,28 > gpurun_out/tma_tensor_n30.jsonl 2>&1
echo "tma_tensor rc=$?"; cut -c1-60,260-400 gpurun_out/tma_tensor_n30.jsonl | tail -8
timeout 20 python -m quantum_simulations_b200.bench.kernel 30 --dtype complex64 --reps 3 --no-observables --json gpurun_out/kernel_table_n30.jsonl > gpurun_out/kernel_table_n30_c64.txt 2>&1
echo "kernel table c64 rc=$?"; tail -3 gpurun_out/kernel_table_n30_c64.txt
