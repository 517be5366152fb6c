#!/bin/bash
# round 2, GPU call D: full parity suite with the new forward kernel / fused scan / late colours, then A/B bench lines
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_d.log; tail -4 gpurun_out/pytest_d.log
for v in v2 v1; do
  export GG_FWD_KERNEL=$v
  python bench.py --steps 60 --warmup 10 --no-cpu-baseline > gpurun_out/bench_d_$v.json 2> gpurun_out/bench_d_$v.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_d_$v.json").read().splitlines() if l.startswith("{")][-1])
    print("fwd $v value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],4), d["cuda_graphs"]["mode"][:40], {k["kernel"]:k["ms"] for k in d["roofline"]["kernels"]}, d["roofline"]["kernel_ms_sum"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_d_$v.err").read()[-3000:])
PY
done
