"""e2e (host buffers in and out) per pipeline depth at configs[1].  Output: one line per depth."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relation_detr_b200 import workloads
from relation_detr_b200.hostpipe import MsdaHostPipeline
shape = workloads.MSDA_SHAPES["msda_enc_800x1333_b8"]
inp = workloads.make_msda_inputs(shape, "S", seed=0, device="cuda:0")
host = {k: inp[k].cpu().pin_memory() for k in ("value", "sampling_locations", "attention_weights", "grad_output")}
for depth in (2, 3, 4, 2, 3):
    pipe = MsdaHostPipeline(inp["spatial_shapes"], inp["level_start_index"], inp["value"].device, depth=depth)
    for _ in range(4):
        pipe.submit(host)
    pipe.wait(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 12
    e0.record(pipe.s_in)
    for _ in range(steps):
        pipe.submit(host)
    e1.record(pipe.s_out)
    pipe.wait(); torch.cuda.synchronize()
    print("depth", depth, "ms/step %.3f" % (e0.elapsed_time(e1) / steps), flush=True)
    del pipe
    torch.cuda.empty_cache()
