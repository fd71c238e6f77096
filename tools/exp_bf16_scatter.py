"""bf16 backward per rdetr_msda_set_bf16_scatter threshold.  One line per (shape, threshold)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys, torch
sys.path.insert(0, %r)
from relation_detr_b200 import ops, workloads
shape = workloads.MSDA_SHAPES[os.environ.get("SHAPE", "msda_enc_800x1333_b8")]
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
res = []
for kind in ("S", "U"):
    inp = workloads.make_msda_inputs(shape, kind, seed=0, device="cuda:0")
    v = inp["value"].bfloat16(); g = inp["grad_output"].bfloat16()
    a = (v, inp["spatial_shapes"], inp["level_start_index"], inp["sampling_locations"], inp["attention_weights"], g)
    res.append("%%s/bf16 bwd %%.4f" %% (kind, t(lambda: ops.msda_backward(*a))))
print(os.environ.get("TAG"), " | ".join(res))
''' % ROOT
for shape in ("msda_enc_800x1333_b8", "msda_enc_800x1333_b2", "msda_enc_1200x2000_b1", "msda_dec_900_b8"):
    for thr in ("0", "32", "100", "400", "100000"):
        env = dict(os.environ, TAG="%s bf16_scatter=%s" % (shape, thr), SHAPE=shape, RDETR_MSDA_BF16_SCATTER=thr)
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=200)
        print(out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:], flush=True)
