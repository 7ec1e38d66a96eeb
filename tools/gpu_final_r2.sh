#!/bin/bash
# Final single-GPU validation of round 2: whole GPU suite, default bench line, ncu launch list.
cd "$(dirname "$0")/.."
out=gpurun_out; mkdir -p $out
timeout 700 python -m pytest tests -m gpu -q > $out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_gpu_final.log
timeout 400 python bench.py --steps 5 --warmup 3 > $out/bench_n30_final.json 2>$out/bench_n30_final.err; echo "bench rc=$?"; python tools/show_bench.py $out/bench_n30_final.json | tail -3
timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-zero-support --no-others > $out/plain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/ncu_launches_final.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-zero-support --no-others > $out/ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke rc=$?"
