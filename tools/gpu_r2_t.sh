#!/bin/bash
# 8 GPUs: final bench line (no L2 flush), exchange kernel with 8 / 16 reductions in flight per thread
set -u
N=8
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
run_bench () {
  tag=$1; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $N > gpurun_out/bench_t_${N}gpu_${tag}.json 2> gpurun_out/bench_t_${N}gpu_${tag}.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_t_${N}gpu_${tag}.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("N=$N $tag value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],4), d["step_ms"], d["e2e_host"], d["cuda_graphs"].get("colour_gate_timed_out"))
    c=d.get("collective") or {}
    print("   ", {k:c.get(k) for k in ("impl","ms_unoverlapped","busbw_gbs")}, (c.get("allreduce_check") or {}).get("ok"))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_t_${N}gpu_${tag}.err").read()[-2000:])
PY
}
run_bench unroll8 GG_AR_UNROLL=8
run_bench unroll16 GG_AR_UNROLL=16
