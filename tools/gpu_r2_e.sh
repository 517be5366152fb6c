#!/bin/bash
# round 2, GPU call E: ncu full capture of one launch of every kernel of the step (eager bench, after warm-up)
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"_kernel" -s 80 -c 14 -o gpurun_out/r2_prof python bench.py --eager --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"
ls -la gpurun_out | tail -3
