#!/bin/bash
# Round-2 single-GPU validation: whole GPU suite, full-size parity (n = 30 vs the C oracle), default bench line,
# ncu launch list and one full capture of the pass kernels of one circuit execution.
cd "$(dirname "$0")/.."
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -x > $out/pytest_gpu_r2_full.log 2>&1; echo "pytest rc=$?"; tail -4 $out/pytest_gpu_r2_full.log
QSV_TEST_FULL=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "n30_full_vector" > $out/parity_n30_full_vector.log 2>&1; echo "n30 rc=$?"; grep -E "max\|d\||passed|failed" $out/parity_n30_full_vector.log | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 > $out/bench_n30_r2.json 2>$out/bench_n30_r2.err; echo "bench rc=$?"; python tools/show_bench.py $out/bench_n30_r2.json | tail -3
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-zero-support --no-others > $out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/ncu_launches_r02.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-zero-support --no-others > $out/ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pass_jit -s 22 -c 7 -f -o $out/prof_r02_pass \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-zero-support --no-others > $out/ncu_full.log 2>&1; echo "ncu full rc=$?"
