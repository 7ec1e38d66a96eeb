#!/bin/bash
cd "$(dirname "$0")/.."
out=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name k_pass_ring --launch-skip 9 --launch-count 1 -o $out/prof_r01d -f python tools/ncu_cases.py 28 > $out/ncu_r01d.log 2>&1; echo "ncu rc=$?"; tail -2 $out/ncu_r01d.log
