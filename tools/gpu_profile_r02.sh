#!/bin/bash
# Round-2 refresh of the profiler evidence bench.py quotes: ncu launch list of `bench.py --quick` (kernel shares of the step)
# and one `ncu --set full` capture of the MSDA kernels at configs[1] (roofline.traffic via tools/record_traffic.py).
set -u
mkdir -p gpurun_out
TAG=${1:-r02x}
timeout 600 python bench.py --steps 3 --warmup 3 --quick --no-cpu-baseline --no-train > gpurun_out/plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 3 --warmup 3 --quick --no-cpu-baseline --no-train > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
timeout 300 python tools/profile_ops.py msda --iters 2 > gpurun_out/plain_profile.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'msda_' -c 4 -f -o gpurun_out/prof_${TAG}_msda python tools/profile_ops.py msda --iters 2 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
