#!/bin/bash
# 1 GPU: forked colour kernel grid variants
set -u
mkdir -p gpurun_out
for v in "GG_FWD_FORK=0" "GG_SH_FORK_BLOCKS=100000" "GG_SH_FORK_BLOCKS=296" "GG_SH_FORK_BLOCKS=592" "GG_SH_FORK_BLOCKS=888"; do
env $v timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_n.json 2> gpurun_out/bench_n.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_n.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("$v value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],4), d["step_ms"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_n.err").read()[-2000:])
PY
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
