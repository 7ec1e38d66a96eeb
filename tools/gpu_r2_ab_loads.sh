#!/bin/bash
# Round 2, last GPU call: A/B of the pass kernel's data-movement switches on the headline workload (n=30 complex128),
# parity of the fastest variant, and the tile-shape sweep of the copy skeleton (tools/tma_tensor.cu).
#   gpurun --timeout 110 -- bash tools/gpu_r2_ab_loads.sh
mkdir -p gpurun_out/ab
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-zero-support --no-others"
for cfg in "0 0" "4 0" "1 1" "3 1" "2 2" "4 2"; do
  set -- $cfg
  QSV_JIT_TILE_BLOCK=$1 QSV_JIT_PAIR=$2 timeout 20 $B > gpurun_out/ab/bench_blk$1_pair$2.json 2> gpurun_out/ab/bench_blk$1_pair$2.err
  echo "blk=$1 pair=$2 rc=$?"
done
python tools/ab_pick.py gpurun_out/ab > gpurun_out/ab/pick.env
cat gpurun_out/ab/pick.env
source gpurun_out/ab/pick.env
timeout 40 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "n26 or replay" > gpurun_out/ab/pytest_parity_picked.log 2>&1
echo "parity (blk=$QSV_JIT_TILE_BLOCK pair=$QSV_JIT_PAIR) rc=$?"; tail -2 gpurun_out/ab/pytest_parity_picked.log
timeout 20 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
TMA_ONLY=cp TMA_BLK=0,2 TMA_PAIR=0,1,2 timeout 25 tools/_build/tma_tensor 30 3 6 \
  3,4,5,6,7,8,9,10 22,23,24,25,26,27,28,29 11,12,13,14,15,16,17,18 11,19,20,21,22,23,27,29 4,7,14,17,25,27,28,29 9,10,11,12,20,24,25,26 \
  3,5,6,8,10,15,17,18 3,5,8,15,16,18,19,22 3,4,5,13,14,17,18,19 \
  3,23,24,25,26,27,28,29 4,23,24,25,26,27,28,29 7,23,24,25,26,27,28,29 3,4,24,25,26,27,28,29 3,5,24,25,26,27,28,29 4,5,24,25,26,27,28,29 \
  3,4,5,25,26,27,28,29 5,6,7,25,26,27,28,29 3,4,5,6,26,27,28,29 5,6,7,8,9,10,11,12 7,8,9,10,11,12,13,14 15,16,17,18,19,20,21,22 \
  3,11,12,13,14,15,16,17 3,4,11,12,13,14,15,16 > gpurun_out/ab/tma_tensor_sweep_n30.jsonl 2>&1
echo "sweep rc=$?"; wc -l gpurun_out/ab/tma_tensor_sweep_n30.jsonl
