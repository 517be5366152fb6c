#!/bin/bash
# round 2, GPU call C (N GPUs): exchange probe (NCCL vs in-switch multimem kernel) + multi-GPU bench lines (graph / eager)
set -u
N=${1:-2}
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/allreduce_probe.py > gpurun_out/probe_${N}gpu.json 2> gpurun_out/probe_${N}gpu.err
grep '^{' gpurun_out/probe_${N}gpu.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for k in ('nccl','nvls_multimem'):
    if k in d: print(k, {x:(round(v,4) if isinstance(v,float) else v) for x,v in d[k].items() if x!='check'}, d[k]['check']['ok'], d[k]['check']['rel_err'])
"
tail -3 gpurun_out/probe_${N}gpu.err
for m in "" "--eager"; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 40 --warmup 10 $m > gpurun_out/bench_c_${N}gpu$m.json 2> gpurun_out/bench_c_${N}gpu$m.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_c_${N}gpu$m.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("N=$N $m value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],4), "host_ms", round(d.get("host_ms_per_step",0),3), d["cuda_graphs"]["mode"][:60], "h2d GB/s", round(d["e2e"].get("h2d_gbs_measured",0),1))
    c=d.get("collective") or {}
    print({k:c.get(k) for k in ("impl","ms_unoverlapped","busbw_gbs")}, (c.get("allreduce_check") or {}).get("ok"))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_c_${N}gpu$m.err").read()[-3000:])
PY
done
