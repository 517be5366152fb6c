#!/bin/bash
# 2 GPUs: the bench line through the default path (checks the end-to-end loop with exchanges)
set -u
N=${1:-2}
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N > gpurun_out/bench_s_${N}gpu.json 2> gpurun_out/bench_s_${N}gpu.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_s_${N}gpu.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("N=$N value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],4), d["step_ms"], d["e2e_host"], d["cuda_graphs"].get("colour_gate_timed_out"), d["cuda_graphs"].get("exchange"))
    c=d.get("collective") or {}
    print("   ", {k:c.get(k) for k in ("impl","ms_unoverlapped","busbw_gbs")}, (c.get("allreduce_check") or {}).get("ok"))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_s_${N}gpu.err").read()[-2000:])
PY
