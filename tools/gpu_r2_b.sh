#!/bin/bash
# round 2, GPU call B: full parity suite (facade replay, N1 variants, N3) + one bench line
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s 2>&1 | tail -80 > gpurun_out/pytest_b.log
tail -15 gpurun_out/pytest_b.log
python bench.py --steps 40 --warmup 8 --no-cpu-baseline > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_b.json").read().strip().splitlines()[-1])
    ks={k["kernel"]:k["ms"] for k in d["roofline"]["kernels"]}
    print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],4), "ksum", d["roofline"]["kernel_ms_sum"], ks)
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_b.err").read()[-3000:])
PY
