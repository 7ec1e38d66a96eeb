#!/bin/bash
# second GPU call: TMA ring microbenchmark sweep + ncu capture of the v1 pass kernel
cd "$(dirname "$0")/.."
out=gpurun_out
B=tools/_build/tma_stream
: > $out/tma_stream.jsonl
run() { timeout 60 $B "$@" >> $out/tma_stream.jsonl 2>&1; }
N=30
# t=11: row size, ring depth, groups
for a in 3 5 7; do for nbuf in 4 6 7; do for g in 2 4; do run $N 11 $a $nbuf $g 0 0; done; done; done
# exchange rounds and compute delay (t=11, a=5, nbuf=7)
for g in 2 3 4; do for xch in 1 2 4 6; do run $N 11 5 7 $g 0 $xch; done; done
for g in 2 4; do for d in 1000 2000 4000 8000; do run $N 11 5 7 $g $d 0; done; done
for g in 4; do for d in 2000 4000; do for xch in 2 4; do run $N 11 5 7 $g $d $xch; done; done; done
# t=12 (64 KB tiles): 3 buffers
for a in 5 7; do for g in 1 2; do for xch in 0 2 4; do run $N 12 $a 3 $g 0 $xch; done; done; done
# t=10
for g in 4; do for nbuf in 8 14; do run $N 10 5 $nbuf $g 0 0; run $N 10 5 $nbuf $g 0 3; done; done
# 2 CTAs per SM, t=11, nbuf 3
for g in 2; do for xch in 0 2 4; do run $N 11 5 3 $g 0 $xch 2; done; done
echo "tma sweep done: $(wc -l < $out/tma_stream.jsonl) lines"
python tools/ncu_cases.py 28 > $out/ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_pass -s 5 -c 5 -f -o $out/prof_v1 python tools/ncu_cases.py 28 > $out/ncu_run.log 2>&1
echo "ncu rc=$?"
tail -3 $out/ncu_run.log
