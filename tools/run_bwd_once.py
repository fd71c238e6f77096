import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relation_detr_b200 import ops, workloads
shape = workloads.MSDA_SHAPES[os.environ.get("SHAPE", "msda_enc_800x1333_b8")]
inp = workloads.make_msda_inputs(shape, os.environ.get("LOC", "S"), seed=0, device="cuda:0")
a = (inp["value"], inp["spatial_shapes"], inp["level_start_index"], inp["sampling_locations"], inp["attention_weights"], inp["grad_output"])
for _ in range(int(os.environ.get("N", "2"))):
    ops.msda_backward(*a)
torch.cuda.synchronize()
print("ok")
