#!/bin/bash
# One gpurun call: GPU tests, bench, then ncu (launch list + one full capture of the MSDA kernels).
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q -s --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== ncu launch list (bench --quick)"
timeout 600 python bench.py --steps 3 --warmup 3 --quick --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 3 --warmup 3 --quick --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
echo "== ncu full (msda + rel kernels)"
timeout 300 python tools/profile_ops.py all --iters 2 > gpurun_out/plain_profile.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'msda_|rel_' -c 12 -f -o gpurun_out/prof_${TAG} python tools/profile_ops.py all --iters 2 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
