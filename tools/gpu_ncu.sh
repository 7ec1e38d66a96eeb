#!/bin/bash
# ncu evidence for the bench workload: launch list (timing only) + one full capture of the 8 pass kernels
cd "$(dirname "$0")/.."
out=gpurun_out
python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > $out/bench_pre_ncu.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_r01.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > $out/ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name k_pass_jit --launch-skip 8 --launch-count 8 -o $out/prof_r01_jit -f python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > $out/ncu_jit.log 2>&1; echo "ncu full rc=$?"
