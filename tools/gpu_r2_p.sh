#!/bin/bash
# 1 GPU: racecheck over the barrier-synchronised v1 kernels, v1-vs-v2 equivalence test, bench
set -u
mkdir -p gpurun_out
SEL_RACE="kat_single or depth_ties or late_colour or lazy_forward_path_bucketed or registration_style or photometric_l1_with_8bit"
GG_FWD_KERNEL=v1 GG_BWD_PATH=v1 timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 99 --log-file gpurun_out/r2_racecheck_v1.log python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL_RACE" > gpurun_out/r2_racecheck_v1_pytest.txt 2>&1
echo "racecheck v1 rc=$?"; tail -2 gpurun_out/r2_racecheck_v1_pytest.txt; grep "RACECHECK SUMMARY" gpurun_out/r2_racecheck_v1.log | tail -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_p_1gpu.json 2> gpurun_out/bench_p_1gpu.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_p_1gpu.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("N=1 value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],4), d["step_ms"])
    print({k["kernel"]:k["ms"] for k in d["roofline"]["kernels"]})
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_p_1gpu.err").read()[-3000:])
PY
