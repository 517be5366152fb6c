#!/bin/bash
# 8 GPUs: final bench line + two exchange-concurrency variants
set -u
N=8
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
run_bench () {
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/bench_q_${N}gpu_${tag}.json 2> gpurun_out/bench_q_${N}gpu_${tag}.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_q_${N}gpu_${tag}.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("N=$N $tag value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],4), d["step_ms"], d["e2e_host"], d["cuda_graphs"].get("colour_gate_timed_out"))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_q_${N}gpu_${tag}.err").read()[-2000:])
PY
}
run_bench default GG_BENCH_EXCHANGE=around
run_bench shfirst_43_21 GG_BENCH_SH_FIRST=1 GG_AR_BLOCKS=43,21
run_bench shfirst_52_12 GG_BENCH_SH_FIRST=1 GG_AR_BLOCKS=52,12
