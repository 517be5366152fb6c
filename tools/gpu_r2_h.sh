#!/bin/bash
# N-GPU: exchange-kernel grid sweep (probe), then the bench line with the default
set -u
N=${1:-8}
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
for b in "32,64" "64,128" "128,256"; do
export GG_AR_BLOCKS=$b
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/allreduce_probe.py 2> gpurun_out/probe_h.err | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for k in ('nccl','nvls_multimem'):
    if k in d: print('blocks $b', k, {x:(round(v,4) if isinstance(v,float) else v) for x,v in d[k].items() if x in ('full_ms','geometry_ms','sh_ms','full_busbw_gbs')}, d[k]['check']['ok'])
"
done
unset GG_AR_BLOCKS
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 40 --warmup 10 > gpurun_out/bench_h_${N}gpu.json 2> gpurun_out/bench_h_${N}gpu.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_h_${N}gpu.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("N=$N value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],4), d["step_ms"], d["cuda_graphs"]["mode"][:40])
    c=d.get("collective") or {}
    print({k:c.get(k) for k in ("impl","ms_unoverlapped","busbw_gbs")}, (c.get("allreduce_check") or {}).get("ok"))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_h_${N}gpu.err").read()[-3000:])
PY
