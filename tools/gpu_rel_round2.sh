#!/bin/bash
# REL FAST kernels (round 2, packed fp32x2): parity, timing (4 / 8 rows per iteration), then one ncu --set full capture.
set -u
mkdir -p gpurun_out
T=${1:-r02r}
python -m pytest tests/test_rel_gpu.py tests/test_rel_attention_gpu.py -q > gpurun_out/${T}_pytest_rel.txt 2>&1; tail -4 gpurun_out/${T}_pytest_rel.txt
python tools/time_rel.py > gpurun_out/${T}_time_rel_rows4.json 2>&1; cat gpurun_out/${T}_time_rel_rows4.json
RDETR_REL_ROWS=8 python tools/time_rel.py > gpurun_out/${T}_time_rel_rows8.json 2>&1; cat gpurun_out/${T}_time_rel_rows8.json
python tools/time_rel_attention.py > gpurun_out/${T}_time_rel_attention.jsonl 2>&1; cat gpurun_out/${T}_time_rel_attention.jsonl
timeout 300 python tools/profile_ops.py rel --iters 1 > gpurun_out/plain_profile.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'rel_(fwd|bwd)_fast' -c 2 -f -o gpurun_out/prof_${T}_rel python tools/profile_ops.py rel --iters 1 > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"
