"""Training-step proxy for BASELINE.json configs[3] (R50 800x1333, batch 2 per GPU, DDP at 1/2/4/8 GPUs).

The reference's detector (backbone, matcher, losses) is outside this repository's scope and its code does
not exist on the GPU box, so this proxy runs what the hot path lives in: a deformable encoder (6 layers,
MSDA self-attention over S = 22323 tokens) and the relation decoder of tools/decoder_harness.py (main pass
with position-relation bias + CDN mask, hybrid pass without), forward + backward + AdamW step, under
torch DDP over NCCL (gradient all-reduce is the only collective).  Inputs are synthetic multi-level
features (what backbone + neck would produce).  Reports images/s = world * batch / step time
(CUDA events, MAX over ranks).

    python tools/train_proxy.py [--bf16] [--steps K]          # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/train_proxy.py ...
"""
import argparse
import json
import os
import sys

import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import decoder_harness as dh  # noqa: E402
import relation_detr_b200 as rd  # noqa: E402
from relation_detr_b200 import dist as rdist  # noqa: E402
from relation_detr_b200 import workloads  # noqa: E402


class EncoderLayer(nn.Module):
    """Call pattern of RelationTransformerEncoderLayer (relation_transformer.py:208-276): MSDA self-attention
    with query = x + pos, value = x, 2-d reference points on the pixel grid; then FFN."""

    def __init__(self, d=256, ffn=1024, heads=8, levels=4, points=4):
        super().__init__()
        self.self_attn = rd.MultiScaleDeformableAttention(d, levels, heads, points)
        self.norm1, self.norm2 = nn.LayerNorm(d), nn.LayerNorm(d)
        self.linear1, self.linear2 = nn.Linear(d, ffn), nn.Linear(ffn, d)

    def forward(self, x, pos, ref, ss, lsi, mask):
        x = self.norm1(x + self.self_attn(x + pos, ref, x, ss, lsi, mask))
        return self.norm2(x + self.linear2(torch.relu(self.linear1(x))))


class BackboneNeck(nn.Module):
    """What feeds the transformer in configs[3]: torchvision ResNet-50 (random init, frozen BN, stem + layer1
    frozen as upstream does) -> C3..C5 -> 1x1 conv + GroupNorm to 256 channels, plus a stride-2 3x3 conv level
    (call pattern of models/necks/channel_mapper.py).  Library code (cuDNN), not part of the hot path."""

    def __init__(self, d=256):
        super().__init__()
        import torchvision
        from torchvision.ops.misc import FrozenBatchNorm2d

        r = torchvision.models.resnet50(weights=None, norm_layer=FrozenBatchNorm2d)
        self.stem = nn.Sequential(r.conv1, r.bn1, r.relu, r.maxpool, r.layer1)
        for p in self.stem.parameters():
            p.requires_grad_(False)
        self.layer2, self.layer3, self.layer4 = r.layer2, r.layer3, r.layer4
        self.lateral = nn.ModuleList([nn.Sequential(nn.Conv2d(c, d, 1), nn.GroupNorm(32, d)) for c in (512, 1024, 2048)])
        self.extra = nn.Sequential(nn.Conv2d(2048, d, 3, stride=2, padding=1), nn.GroupNorm(32, d))

    def forward(self, images):
        with torch.no_grad():
            x = self.stem(images)
        c3 = self.layer2(x)
        c4 = self.layer3(c3)
        c5 = self.layer4(c4)
        feats = [lat(c) for lat, c in zip(self.lateral, (c3, c4, c5))] + [self.extra(c5)]
        return torch.cat([f.flatten(2).transpose(1, 2) for f in feats], 1)  # [B, S, 256]


class TransformerProxy(nn.Module):
    def __init__(self, levels=4, backbone=False):
        super().__init__()
        self.backbone = BackboneNeck() if backbone else None
        self.encoder = nn.ModuleList([EncoderLayer(levels=levels) for _ in range(6)])
        self.decoder = dh.RelationDecoder("ours", levels=levels)  # shared by the main and the hybrid pass, as upstream

    def forward(self, feats, pos, ref2d, ss, lsi, main, hybrid):
        x = self.backbone(feats) if self.backbone is not None else feats  # feats = images when the backbone is on
        for layer in self.encoder:
            x = layer(x, pos, ref2d, ss, lsi, None)
        c1, b1 = self.decoder(main["query"], main["reference_points"], x, ss, lsi, main["valid_ratios"], main["attn_mask"])
        c2, b2 = self.decoder(hybrid["query"], hybrid["reference_points"], x, ss, lsi, hybrid["valid_ratios"], None,
                                     skip_relation=True)
        loss = c1.square().mean() + b1.square().mean() + c2.square().mean() + b2.square().mean()
        return loss, c1, b1, c2, b2


def matched_box_loss(how, matcher, c1, b1, c2, b2, gts, n_dn, hybrid_assign=6):
    """Bipartite matching of every decoder layer's predictions (main pass: the 900 matching queries after the
    denoising rows; hybrid pass: 1500 queries against ``hybrid_assign`` copies of the targets,
    relation_detr.py:96-134) followed by an L1 loss on the matched boxes.  ``how`` = "device": this repository's
    matcher, all problems in two launches, no synchronisation; "scipy": the reference's way, per problem
    eager cost -> ``.cpu()`` -> SciPy (hungarian_matcher.py:74-81)."""
    from scipy.optimize import linear_sum_assignment

    B = c1.shape[1]
    problems = []  # (pred_boxes with grad, detached boxes, logits, gt boxes, gt labels)
    for c, b, skip, rep in ((c1, b1, n_dn, 1), (c2, b2, 0, hybrid_assign)):
        for layer in range(c.shape[0]):
            for i in range(B):
                problems.append((b[layer, i, skip:], b[layer, i, skip:].detach().float(), c[layer, i, skip:].detach().float(),
                                 gts[i][0].repeat(rep, 1), gts[i][1].repeat(rep)))
    if how == "device":
        pairs = matcher.match_batch(*[[p[k] for p in problems] for k in (1, 2, 3, 4)])
    else:
        pairs = []
        for _, pb, pl, gb, gl in problems:
            r, c_ = linear_sum_assignment(matcher.calculate_cost(pb, pl, gb, gl).cpu())
            pairs.append((torch.as_tensor(r), torch.as_tensor(c_)))
    loss = 0.0
    for (pred, _, _, gb, _), (src, tgt) in zip(problems, pairs):
        loss = loss + (pred[src].float() - gb[tgt]).abs().sum()
    return loss / max(1, sum(len(p[3]) for p in problems))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--bf16", action="store_true")
    ap.add_argument("--tf32", action="store_true")
    ap.add_argument("--backbone", action="store_true", help="prepend ResNet-50 + channel mapper on 800x1344 images")
    ap.add_argument("--graph", action="store_true", help="capture the whole step (fwd+bwd+clip+AdamW) in one CUDA graph")
    ap.add_argument("--matching", choices=("none", "device", "scipy"), default="none",
                    help="add per-layer bipartite matching + matched box loss: this repository's device matcher, or the "
                         "reference's cost -> .cpu() -> SciPy")
    ap.add_argument("--gt", type=int, nargs="*", default=[7, 15], help="ground-truth boxes per image (cycled)")
    args = ap.parse_args()
    assert not (args.graph and args.matching != "none"), "--graph is measured without matching"
    rank, local_rank, world = rdist.env_rank_world()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        rdist.init_process_group("nccl")
    torch.backends.cuda.matmul.allow_tf32 = args.tf32
    torch.manual_seed(0)
    model = TransformerProxy(backbone=args.backbone).to(dev)
    nparams = sum(p.numel() for p in model.parameters())
    ddp = nn.parallel.DistributedDataParallel(model, device_ids=[local_rank], find_unused_parameters=False) if world > 1 else model
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, weight_decay=1e-4, capturable=args.graph)
    levels = workloads.LEVELS_800_1333
    ss, lsi = workloads.shape_tensors(levels, dev)
    S = int(ss.prod(1).sum())
    g = torch.Generator(device=dev).manual_seed(100 + rank)  # each rank owns its own images
    feats = (torch.randn((args.batch, 3, 800, 1344), device=dev, generator=g) if args.backbone
             else torch.randn((args.batch, S, 256), device=dev, generator=g))
    pos = torch.randn((args.batch, S, 256), device=dev, generator=g)
    ref2d = workloads.full_reference_points(levels, dev)[None, :, None, :].expand(args.batch, -1, 4, -1).contiguous()
    main_in = dh.make_inputs(args.batch, 900, 200, levels, seed=rank, device=dev)
    hyb_in = dh.make_inputs(args.batch, 1500, 0, levels, seed=50 + rank, device=dev)

    gts = []
    for i in range(args.batch):
        n = args.gt[i % len(args.gt)]
        gts.append((torch.cat([torch.rand((n, 2), device=dev, generator=g) * 0.8 + 0.1, torch.rand((n, 2), device=dev, generator=g) * 0.3 + 0.02], -1),
                    torch.randint(0, 91, (n,), device=dev, generator=g)))
    matcher = rd.HungarianMatcher(cost_class=2, cost_bbox=5, cost_giou=2, fused_cost=args.matching == "device")

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=args.bf16):
            loss, c1, b1, c2, b2 = ddp(feats, pos, ref2d, ss, lsi, main_in, hyb_in)
        if args.matching != "none":
            loss = loss + matched_box_loss(args.matching, matcher, c1, b1, c2, b2, gts, 200)
        loss.backward()
        torch.nn.utils.clip_grad_norm_([p for p in model.parameters() if p.requires_grad], 0.1)
        opt.step()
        return loss

    if args.graph:
        # the operators never synchronise or allocate behind torch's back, so the whole training step is
        # one CUDA graph (the reference's module cannot be captured: host-sync assert, ms_deform_attn.py:313)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(11 if world > 1 else 3):  # DDP needs 11 eager iterations before capture
                step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(graph):
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=args.bf16):
                static_loss = ddp(feats, pos, ref2d, ss, lsi, main_in, hyb_in)[0]
            static_loss.backward()
            torch.nn.utils.clip_grad_norm_([p for p in model.parameters() if p.requires_grad], 0.1)
            opt.step()

        def step():  # noqa: F811
            graph.replay()
            return static_loss

    for _ in range(args.warmup):
        step()
    rdist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    rdist.barrier()
    ms = rdist.max_over_ranks(e0.elapsed_time(e1) / args.steps, dev)
    if rank == 0:
        print(json.dumps({"workload": ("train_proxy: " + ("ResNet-50 + channel mapper (800x1344 images) + " if args.backbone else "")
                                       + "6-layer deformable encoder + relation decoder (main + hybrid), fwd+bwd+AdamW, DDP/NCCL"),
                          "n_gpus": world, "batch_per_gpu": args.batch, "precision": "bf16 autocast" if args.bf16 else ("tf32 matmul" if args.tf32 else "fp32"),
                          "cuda_graph": bool(args.graph), "matching": args.matching, "params_M": round(nparams / 1e6, 2), "ms_per_step": round(ms, 3),
                          "imgs_per_s": round(world * args.batch / ms * 1e3, 2), "loss": float(loss.detach())}))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
