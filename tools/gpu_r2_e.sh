#!/bin/bash
# round 2, GPU call E: parity suite + ncu (launch list, full capture of one launch of every kernel)
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_e.log; tail -3 gpurun_out/pytest_e.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches.csv python bench.py --eager --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gg::" -s 60 -c 13 -o gpurun_out/r2_prof python bench.py --eager --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"
ls -la gpurun_out | tail -5
