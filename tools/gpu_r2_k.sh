#!/bin/bash
# 1 GPU: full GPU suite + bench with the register-shuffle sort
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/k_pytest.txt
tail -4 gpurun_out/k_pytest.txt
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_k_1gpu.json 2> gpurun_out/bench_k_1gpu.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_k_1gpu.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("N=1 value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],4), "host", d["host_ms_per_step"])
    print({k["kernel"]:k["ms"] for k in d["roofline"]["kernels"]})
    print(d["with_ssim_loss"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_k_1gpu.err").read()[-3000:])
PY
