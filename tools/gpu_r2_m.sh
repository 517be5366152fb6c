#!/bin/bash
# 4 GPUs: bench at N=4 and N=2 (NUMA binding on / off at N=4), cfg4 with the fused binding, probe
set -u
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
run_bench () {  # N, tag, extra env
  N=$1; tag=$2; shift; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/bench_m_${N}gpu_${tag}.json 2> gpurun_out/bench_m_${N}gpu_${tag}.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_m_${N}gpu_${tag}.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("N=$N $tag value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],4), d["e2e_host"], d["e2e"]["h2d_gbs_min_over_ranks"], d["e2e"]["host_numa_binding"])
    c=d.get("collective") or {}
    print("   ", {k:c.get(k) for k in ("impl","ms_unoverlapped","busbw_gbs")}, (c.get("allreduce_check") or {}).get("ok"))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_m_${N}gpu_${tag}.err").read()[-2000:])
PY
}
run_bench 4 numa GG_BENCH_NUMA=1
run_bench 4 nonuma GG_BENCH_NUMA=0
run_bench 2 numa GG_BENCH_NUMA=1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 tools/allreduce_probe.py 2> gpurun_out/probe_m_4gpu.err | grep '^{' > gpurun_out/probe_m_4gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543 tools/run_configs.py --config cfg4 --fused-binding --steps 5 --warmup 2 2> gpurun_out/cfg4_m.err | grep '^{' | tee gpurun_out/cfg4_m_4gpu_fused.json | cut -c1-900
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 tools/run_configs.py --config cfg4 --steps 3 --warmup 1 2> gpurun_out/cfg4_m2.err | grep '^{' | tee gpurun_out/cfg4_m_4gpu_torchchain.json | cut -c1-400
timeout 300 python tools/run_configs.py --config cfg1 --steps 50 --warmup 5 2>/dev/null | grep '^{' | tee gpurun_out/cfg1_m.json | cut -c1-500
nvidia-smi topo -m 2>/dev/null | head -12
