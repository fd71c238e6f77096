"""torch.profiler kernel table for the MSDA *module* (prologue + kernel + projections) at the encoder shape."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import relation_detr_b200 as rd  # noqa: E402
from relation_detr_b200 import workloads  # noqa: E402

dev = "cuda:0"
shape = workloads.MSDA_SHAPES["msda_enc_800x1333_b8"]
torch.manual_seed(0)
mod = rd.MultiScaleDeformableAttention(256, 4, 8, 4).to(dev)
with torch.no_grad():
    mod.sampling_offsets.weight.normal_(0, 0.01)
ss, lsi = workloads.shape_tensors(shape.levels, dev)
x = torch.randn((shape.batch, shape.S, 256), device=dev, requires_grad=True)
pos = torch.randn((shape.batch, shape.S, 256), device=dev)
ref = workloads.full_reference_points(shape.levels, dev)[None, :, None, :].expand(shape.batch, -1, 4, -1).contiguous()


def step():
    out = mod(x + pos, ref, x, ss, lsi, None)
    out.sum().backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=90))
