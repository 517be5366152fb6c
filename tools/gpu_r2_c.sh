#!/bin/bash
# round 2, GPU call C (N GPUs): exchange probe (NCCL vs in-switch multimem kernel) + multi-GPU bench line
set -u
N=${1:-2}
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/allreduce_probe.py > gpurun_out/probe_${N}gpu.json 2> gpurun_out/probe_${N}gpu.err
tail -3 gpurun_out/probe_${N}gpu.json; tail -5 gpurun_out/probe_${N}gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 40 --warmup 8 > gpurun_out/bench_c_${N}gpu.json 2> gpurun_out/bench_c_${N}gpu.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_c_${N}gpu.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("N=$N value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],4), "host_ms", round(d.get("host_ms_per_step",0),3))
    print(json.dumps(d.get("collective"), indent=0)[:1500])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_c_${N}gpu.err").read()[-3000:])
PY
