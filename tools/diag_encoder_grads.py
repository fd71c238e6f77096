"""Where do the gradient differences of the 3-layer encoder integration test come from?  (three numbers per tensor.)
Runs the unmodified RelationTransformerEncoder in fp32 and in fp64 and the install()ed one (fp32), twice each, and prints per
parameter: ours-vs-ref32, ours-vs-ref64, ref32-vs-ref64 (the reference's own fp32 noise), ours-vs-ours (atomics).
python tools/diag_encoder_grads.py  -> JSON lines (needs baseline/_ref and a GPU)."""
import copy
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from baseline import refmodel  # noqa: E402
from relation_detr_b200 import install as rinstall, workloads  # noqa: E402

DEV = "cuda:0"
LEVELS = ((25, 34), (13, 17), (7, 9), (4, 5))


def main():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    refmodel.activate()

    def build():
        from models.bricks import relation_transformer as rt
        layer = rt.RelationTransformerEncoderLayer(embed_dim=256, d_ffn=512, dropout=0.0, n_heads=8,
                                                   activation=torch.nn.ReLU(inplace=True), n_levels=4, n_points=4)
        return rt.RelationTransformerEncoder(layer, num_layers=3)

    rinstall.uninstall()
    torch.manual_seed(0)
    ref = build().to(DEV)
    rinstall.install(fused_memory=True)
    torch.manual_seed(0)
    ours = build().to(DEV)
    rinstall.uninstall()
    with torch.no_grad():
        torch.manual_seed(1)
        for layer in ref.layers:
            layer.self_attn.sampling_offsets.weight.normal_(0, 0.02)
            layer.self_attn.sampling_offsets.bias.add_(torch.randn_like(layer.self_attn.sampling_offsets.bias) * 0.3)
            layer.self_attn.attention_weights.weight.normal_(0, 0.05)
    ours.load_state_dict(ref.state_dict(), strict=True)
    ref64 = copy.deepcopy(ref).double()
    ss, lsi = workloads.shape_tensors(LEVELS, DEV)
    S = int(ss.prod(1).sum())
    g = torch.Generator(device=DEV).manual_seed(7)
    query = torch.randn((2, S, 256), device=DEV, generator=g)
    pos = torch.randn((2, S, 256), device=DEV, generator=g) * 0.1
    refp = workloads.full_reference_points(LEVELS, DEV)[None, :, None, :].expand(2, S, 4, 2).contiguous()
    go = torch.randn((2, S, 256), device=DEV, generator=g)

    def run(m, dt):
        m.zero_grad()
        out = m(query.to(dt), ss, lsi, refp.to(dt), query_pos=pos.to(dt), query_key_padding_mask=None)
        out.backward(go.to(dt))
        return {"out": out.detach().double(), **{n: p.grad.detach().double().clone() for n, p in m.named_parameters() if p.grad is not None}}

    r64 = run(ref64, torch.float64)
    r32 = run(ref, torch.float32)
    o1 = run(ours, torch.float32)
    o2 = run(ours, torch.float32)

    def err(a, b):
        return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))

    worst = {}
    for n in r64:
        row = {"name": n, "ours_vs_ref32": err(o1[n], r32[n]), "ours_vs_ref64": err(o1[n], r64[n]), "ref32_vs_ref64": err(r32[n], r64[n]),
               "ours_vs_ours": err(o1[n], o2[n])}
        print(json.dumps({k: (v if isinstance(v, str) else float(f"{v:.3e}")) for k, v in row.items()}), flush=True)
        for k, v in row.items():
            if k != "name":
                worst[k] = max(worst.get(k, 0.0), v)
    print(json.dumps({"worst": {k: float(f"{v:.3e}") for k, v in worst.items()}}))


if __name__ == "__main__":
    main()
