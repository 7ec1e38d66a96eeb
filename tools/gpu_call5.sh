#!/bin/bash
cd "$(dirname "$0")/.."
out=gpurun_out
timeout 120 python tools/ncu_cases.py 28 > $out/ncu_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass -s 5 -c 5 -f -o $out/prof_ring python tools/ncu_cases.py 28 > $out/ncu_run.log 2>&1
echo "ncu rc=$?"; tail -2 $out/ncu_run.log
