#!/bin/bash
# ncu --set full capture of the two matcher kernels (one step's 28 problems), after the same command ran plainly.
set -u
mkdir -p gpurun_out
timeout 200 python tools/time_matcher.py --iters 2 > gpurun_out/plain_matcher.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'lsap_kernel|match_cost_kernel' -c 4 -f -o gpurun_out/prof_matcher python tools/time_matcher.py --iters 1 > gpurun_out/ncu_matcher.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_matcher.log
