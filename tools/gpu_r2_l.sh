#!/bin/bash
# 8 GPUs: exchange placement variants, probe, cfg5, around-check
set -u
N=8
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
run_bench () {  # tag, extra env..., uses $N
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 80 --warmup 10 > gpurun_out/bench_l_${N}gpu_${tag}.json 2> gpurun_out/bench_l_${N}gpu_${tag}.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_l_${N}gpu_${tag}.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("N=$N $tag value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],4), d["step_ms"], d["cuda_graphs"].get("exchange"), d["cuda_graphs"].get("colour_gate_timed_out"))
    c=d.get("collective") or {}
    print("   ", {k:c.get(k) for k in ("impl","ms_unoverlapped","busbw_gbs")}, (c.get("allreduce_check") or {}).get("ok"))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_l_${N}gpu_${tag}.err").read()[-2000:])
PY
}
run_bench around GG_BENCH_EXCHANGE=around
run_bench shfirst GG_BENCH_EXCHANGE=around GG_BENCH_SH_FIRST=1
run_bench ingraph GG_BENCH_EXCHANGE=ingraph
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 tools/allreduce_probe.py 2> gpurun_out/probe_l_8gpu.err | grep '^{' > gpurun_out/probe_l_8gpu.json
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/probe_l_8gpu.json").read())
    for k in ('nccl','nvls_multimem'):
        if k in d: print(k, {x:(round(v,4) if isinstance(v,float) else v) for x,v in d[k].items() if x in ('full_ms','geometry_ms','sh_ms','full_busbw_gbs')}, d[k]['check']['ok'])
except Exception as e:
    print("probe failed", e)
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/exchange_around_check.py 2> gpurun_out/around_check_8gpu.err | grep '^{' | tee gpurun_out/around_check_8gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 tools/run_configs.py --config cfg5 --steps 2 --warmup 1 2> gpurun_out/cfg5_l.err | grep '^{' | tee gpurun_out/cfg5_l_8gpu.json | cut -c1-1200
tail -2 gpurun_out/cfg5_l.err
