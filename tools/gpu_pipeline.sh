#!/bin/bash
# Pipelined stage transitions on N real GPUs: parity (one-process shards on device 0, then one process per
# GPU over CUDA IPC), then bench.py --gpus N with the pipeline on / off / old pair kernel, and the SM split.
N=${1:-2}
cd "$(dirname "$0")/.."
out=gpurun_out; mkdir -p $out
timeout 400 python -m pytest tests/test_pipelined_swap.py -m gpu -x -q > $out/pytest_pipe_1proc.log 2>&1; echo "one-process rc=$?"; tail -3 $out/pytest_pipe_1proc.log
timeout 400 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -s -k "match_oracle and not scatter" > $out/pytest_mgpu_$N.log 2>&1; echo "torchrun parity rc=$?"; grep -E "max\|d\||passed|failed|Error" $out/pytest_mgpu_$N.log | tail -14
run() {  # tag, env
  env $2 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 \
      bench.py --gpus $N --steps 3 --warmup 3 --no-e2e $3 > $out/bench_n${N}_$1.log 2>$out/bench_n${N}_$1.err; echo "bench $1 rc=$?"
  python - $out/bench_n${N}_$1.log <<'PY'
import json, sys
f = sys.argv[1]
try:
    r = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "n", r["config"]["n_qubits"], "ms/step", round(r["ms_per_step"], 2), r["config"]["step_sequence"],
          "swaps", [(s["bits"], s["ms"], s["gbs_per_direction"]) for s in r["nvlink"]["swaps"]], "pipelined", r.get("pipelined_swap", {}).get("regions"),
          "pass avg", round(r["roofline"]["avg_launch_ms"], 2))
except Exception as e:
    print(f, "ERR", e, open(f).read()[-300:], open(f.replace(".log", ".err")).read()[-1800:])
PY
}
run pipe "QSV_PIPELINE=1"
run plain_tma "QSV_PIPELINE=0"
run plain_pairs "QSV_PIPELINE=0 QSV_SWAP_KERNEL=pairs"
run pipe_sms12 "QSV_PIPELINE=1 QSV_XCHG_SMS=12"
run pipe_sms20 "QSV_PIPELINE=1 QSV_XCHG_SMS=20"
