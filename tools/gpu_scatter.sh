#!/bin/bash
# Scatter passes (qsv_pass_scatter, ABI v6) — first hardware run of the path.
#   bash tools/gpu_scatter.sh        1 GPU: shards of one process on device 0 (raw-pointer wiring)
#   bash tools/gpu_scatter.sh N      N GPUs: + IPC wiring under torchrun, bench with and without --fused-exchange
N=${1:-1}
cd "$(dirname "$0")/.."
out=gpurun_out; mkdir -p $out
export QSV_TEST_SCATTER=1
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -s -k "one_process" > $out/pytest_scatter_1proc.log 2>&1; echo "one-process rc=$?"; tail -5 $out/pytest_scatter_1proc.log
[ "$N" -gt 1 ] || exit 0
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -s -k "scatter_passes" > $out/pytest_scatter_$N.log 2>&1; echo "torchrun rc=$?"; grep -E "max\|d\||passed|failed|skipped" $out/pytest_scatter_$N.log | tail -12
for mode in "" "--fused-exchange"; do
  tag=${mode:+_scatter}
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 \
      bench.py --gpus $N --steps 3 --warmup 3 --no-e2e $mode > $out/bench_n$N$tag.log 2>$out/bench_n$N$tag.err; echo "bench $mode rc=$?"
  python - $out/bench_n$N$tag.log <<'PY'
import json, sys
f = sys.argv[1]
try:
    r = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "n", r["config"]["n_qubits"], "ms/step", round(r["ms_per_step"], 2), r["config"]["step_sequence"],
          "swaps", [(s["bits"], s["ms"]) for s in r["nvlink"]["swaps"]], "scatter", r.get("scatter_pass", {}).get("ms"),
          "pass avg", round(r["roofline"]["avg_launch_ms"], 2))
except Exception as e:
    print(f, "ERR", e, open(f).read()[-500:], open(f.replace(".log", ".err")).read()[-2500:])
PY
done
