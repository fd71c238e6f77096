#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/profile_ops.py rel --iters 1 > gpurun_out/plain_profile.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'rel_fwd' -s 1 -c 1 -f -o gpurun_out/prof_relfwd python tools/profile_ops.py rel --iters 1 > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"
