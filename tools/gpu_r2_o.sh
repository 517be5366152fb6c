#!/bin/bash
# 1 GPU: compute-sanitizer memcheck + racecheck over small parity tests that touch every kernel family
set -u
mkdir -p gpurun_out
SEL_MEM="kat_single or depth_ties or registration_style or zero_gaussians or late_colour or hinted_forward or two_backwards or visibility_mask_is_bit or photometric_l1_with_8bit or inplace_mask or fused_mesh_binding_feeds or lazy_forward_path_bucketed or scale_modifier or densification_cycle"
SEL_RACE="kat_single or depth_ties or late_colour or lazy_forward_path_bucketed or registration_style or photometric_l1_with_8bit"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --log-file gpurun_out/r2_memcheck.log python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL_MEM" > gpurun_out/r2_memcheck_pytest.txt 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/r2_memcheck_pytest.txt; grep -c "Invalid\|ERROR SUMMARY" gpurun_out/r2_memcheck.log; grep "ERROR SUMMARY" gpurun_out/r2_memcheck.log | tail -2
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 99 --log-file gpurun_out/r2_racecheck.log python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL_RACE" > gpurun_out/r2_racecheck_pytest.txt 2>&1
echo "racecheck rc=$?"; tail -3 gpurun_out/r2_racecheck_pytest.txt; grep "RACECHECK SUMMARY\|hazard" gpurun_out/r2_racecheck.log | tail -5
# the barrier-synchronised v1 kernels under racecheck (expected clean), and the v1-vs-v2 equivalence test
GG_FWD_KERNEL=v1 GG_BWD_PATH=v1 timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 99 --log-file gpurun_out/r2_racecheck_v1.log python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL_RACE" > gpurun_out/r2_racecheck_v1_pytest.txt 2>&1
echo "racecheck v1 rc=$?"; tail -2 gpurun_out/r2_racecheck_v1_pytest.txt; grep "RACECHECK SUMMARY" gpurun_out/r2_racecheck_v1.log | tail -2
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "decoupled_warp" 2>&1 | tail -3
