#!/bin/bash
# ncu --set full of the bf16 and fused-prologue instantiations of the MSDA kernels (one launch each)
set -u
mkdir -p gpurun_out
timeout 300 python tools/profile_ops.py msda --dtype bf16 --iters 2 > gpurun_out/plain_profile.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'msda_' -s 2 -c 2 -f -o gpurun_out/prof_bf16 python tools/profile_ops.py msda --dtype bf16 --iters 2 > gpurun_out/ncu_bf16.log 2>&1
echo "bf16 rc=$?"
timeout 300 python tools/profile_ops.py msda --fused --iters 2 > gpurun_out/plain_profile2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'msda_' -s 2 -c 2 -f -o gpurun_out/prof_fused python tools/profile_ops.py msda --fused --iters 2 > gpurun_out/ncu_fused.log 2>&1
echo "fused rc=$?"
