#!/bin/bash
# N GPUs: graph + eager exchanges check, then bench with both exchange placements
set -u
N=${1:-2}
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 tools/exchange_around_check.py 2> gpurun_out/around_check_${N}gpu.err | grep '^{' | tee gpurun_out/around_check_${N}gpu.json
tail -3 gpurun_out/around_check_${N}gpu.err
for mode in around ingraph; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 100 --warmup 10 --exchange $mode > gpurun_out/bench_j_${N}gpu_${mode}.json 2> gpurun_out/bench_j_${N}gpu_${mode}.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_j_${N}gpu_${mode}.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("N=$N $mode value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],4), d["step_ms"], d["cuda_graphs"])
    c=d.get("collective") or {}
    print({k:c.get(k) for k in ("impl","ms_unoverlapped","busbw_gbs")}, (c.get("allreduce_check") or {}).get("ok"))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_j_${N}gpu_${mode}.err").read()[-3000:])
PY
done
