#!/bin/bash
# A/B of the per-tile sort (rank-merge vs bitonic), eager per-kernel timings + graph value
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "binning or cfg1_parity_vs_c or dense or lazy or cfg2" 2>&1 | tail -3
for v in rank bitonic; do
  export GG_SORT=$v
  python bench.py --steps 60 --warmup 10 --no-cpu-baseline > gpurun_out/bench_f_$v.json 2> gpurun_out/bench_f_$v.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_f_$v.json").read().splitlines() if l.startswith("{")][-1])
    print("sort $v value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), {k["kernel"]:k["ms"] for k in d["roofline"]["kernels"]}, d["roofline"]["kernel_ms_sum"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_f_$v.err").read()[-3000:])
PY
done
