#!/bin/bash
# round 2, GPU call A: parity suite, blend_bwd v1-vs-v2 A/B, launch list + full ncu capture of blend_bwd2
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60 > gpurun_out/pytest_a.log
echo "pytest rc=$?" >> gpurun_out/pytest_a.log
tail -5 gpurun_out/pytest_a.log
for v in v1 v2m4 v2m3; do
  case $v in
    v1)   export GG_BWD_PATH=v1; unset GG_BWD2_MINB;;
    v2m4) export GG_BWD_PATH=v2; export GG_BWD2_MINB=4;;
    v2m3) export GG_BWD_PATH=v2; export GG_BWD2_MINB=3;;
  esac
  python bench.py --steps 40 --warmup 8 --no-cpu-baseline > gpurun_out/bench_a_$v.json 2> gpurun_out/bench_a_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_a_$v.json").read().strip().splitlines()[-1])
    ks={k["kernel"]:k["ms"] for k in d["roofline"]["kernels"]}
    print("$v", "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],4), "ksum", d["roofline"]["kernel_ms_sum"], ks)
except Exception as e:
    print("$v bench failed", e); print(open("gpurun_out/bench_a_$v.err").read()[-2000:])
PY
done
export GG_BWD_PATH=v2; unset GG_BWD2_MINB
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blend_bwd2|blend_fwd_kernel|sort_pack" -s 6 -c 3 -o gpurun_out/r2a_prof python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"
ls -la gpurun_out
