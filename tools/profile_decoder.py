"""torch.profiler kernel table for one fwd+bwd of the decoder harness (where does configs[2] time go)."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import decoder_harness as dh  # noqa: E402
from relation_detr_b200 import workloads  # noqa: E402

ours, _ = dh.build_pair(0)
main_inp = dh.make_inputs(8, 900, 200, workloads.LEVELS_800_1333, 0)
hyb = dh.make_inputs(8, 1500, 0, workloads.LEVELS_800_1333, 1)
import contextlib
bf16 = "--bf16" in sys.argv
ctx = (lambda: torch.autocast("cuda", dtype=torch.bfloat16)) if bf16 else contextlib.nullcontext
with ctx():
    dh.time_block(ours, main_inp, hyb, warmup=2, iters=1)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        dh.time_block(ours, main_inp, hyb, warmup=0, iters=1)
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
