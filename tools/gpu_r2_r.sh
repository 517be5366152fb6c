#!/bin/bash
# 1 GPU: final verification -- full GPU suite, smoke, bench
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_r_1gpu.json 2> gpurun_out/bench_r_1gpu.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_r_1gpu.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("N=1 value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],4), d["step_ms"], d["e2e_host"], "launches", d["gpu_launches"], "steps", d["steps"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_r_1gpu.err").read()[-3000:])
PY
