#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/profile_ops.py msda --fused --iters 2 > gpurun_out/plain_profile.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'msda_fwd' -s 1 -c 1 -f -o gpurun_out/prof_fused python tools/profile_ops.py msda --fused --iters 2 > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_full.log
