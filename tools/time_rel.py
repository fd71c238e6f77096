"""Quick CUDA-event timing of the relation kernels (fwd/bwd, EXACT/FAST) at B=8, N=900 / 1100."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relation_detr_b200 import ops, workloads  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

res = {}
for name in ("rel_900_b8", "rel_1100_b8"):
    for fast in (False, True):
        res[f"{name}_{'fast' if fast else 'exact'}"] = bench.time_rel(torch, ops, workloads, name, 20, 5, fast)
print(json.dumps(res))
