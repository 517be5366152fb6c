#!/bin/bash
# 1 GPU: full GPU test suite, bench line, ncu launch list + full capture of one step
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/i_pytest.txt
cat gpurun_out/i_pytest.txt | tail -5
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_i_1gpu.json 2> gpurun_out/bench_i_1gpu.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_i_1gpu.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("N=1 value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],4))
    print(d.get("kernel_ms"))
    print(d["roofline"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_i_1gpu.err").read()[-3000:])
PY
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 4 --warmup 3 --eager > gpurun_out/ncu_i_launch.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"_kernel" --launch-skip 120 --launch-count 40 -o gpurun_out/r2_final -f python bench.py --steps 4 --warmup 3 --eager > gpurun_out/ncu_i_full.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
