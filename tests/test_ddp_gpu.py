"""DDP over the drop-in modules: 2 ranks (gloo backend, both on cuda:0 -- NCCL refuses two ranks on one
device and the round-end GPU box has a single GPU), images sharded by rank.  After the gradient
all-reduce every rank must hold the gradient of the concatenated batch computed in one process
("distributed"), with find_unused_parameters=False as the reference configures DDP
(main.py:106, train_config.py:17)."""
import os
import socket
import sys
import tempfile
import traceback

import pytest
import torch
import torch.multiprocessing as mp

TOOLS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools")
sys.path.insert(0, TOOLS)

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(420)]
LEVELS = ((13, 21), (7, 11), (4, 6), (2, 3))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _loss(model, inp):
    c, b = model(**inp)
    return c.square().mean() + b.square().mean()


PER_IMAGE = ("query", "reference_points", "value", "valid_ratios")  # tensors with a leading batch dim


def _shard(full, lo, hi, n):
    return {k: (v[lo:hi].contiguous() if k in PER_IMAGE else v) for k, v in full.items()}


def _worker(rank, world, port, outdir):
    """Writes grads_<rank>.pt on success or error_<rank>.txt on failure; never blocks the parent."""
    try:
        sys.path.insert(0, TOOLS)
        import torch.distributed as dist

        import decoder_harness as dh
        from relation_detr_b200 import dist as rdist

        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK="0")
        torch.cuda.set_device(0)
        import datetime
        dist.init_process_group("gloo", rank=rank, world_size=world, timeout=datetime.timedelta(seconds=120))
        ours, _ = dh.build_pair(0, layers=2)
        full = dh.make_inputs(4, 30, 20, LEVELS, seed=1)  # the same global batch on every rank
        lo, hi = rdist.shard_range(4, rank, world)
        ddp = torch.nn.parallel.DistributedDataParallel(ours, device_ids=[0], find_unused_parameters=False)
        _loss(ddp, _shard(full, lo, hi, 4)).backward()  # DDP averages the per-rank gradients
        torch.cuda.synchronize()
        grads = {n: p.grad.detach().cpu() for n, p in ours.named_parameters() if p.grad is not None}
        torch.save(grads, os.path.join(outdir, f"grads_{rank}.pt"))
        dist.destroy_process_group()
    except Exception:
        with open(os.path.join(outdir, f"error_{rank}.txt"), "w") as f:
            f.write(traceback.format_exc())


def test_two_rank_ddp_matches_single_process_gradients():
    import decoder_harness as dh

    ctx = mp.get_context("spawn")
    port = _free_port()
    with tempfile.TemporaryDirectory() as outdir:
        procs = [ctx.Process(target=_worker, args=(r, 2, port, outdir)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(240)
        hung = [p for p in procs if p.is_alive()]
        for p in hung:
            p.kill()
        errors = [open(os.path.join(outdir, f)).read() for f in sorted(os.listdir(outdir)) if f.startswith("error_")]
        assert not errors, "\n".join(errors)
        assert not hung, "DDP worker(s) did not finish within 240 s"
        g0 = torch.load(os.path.join(outdir, "grads_0.pt"))
        g1 = torch.load(os.path.join(outdir, "grads_1.pt"))
    assert all(torch.equal(g0[n], g1[n]) for n in g0)  # all-reduced: identical on both ranks

    # single process, whole batch: mean over ranks of per-shard mean losses == DDP's averaged gradient
    ours, _ = dh.build_pair(0, layers=2)
    full = dh.make_inputs(4, 30, 20, LEVELS, seed=1)
    total = 0.0
    for lo, hi in ((0, 2), (2, 4)):
        total = total + _loss(ours, _shard(full, lo, hi, 4)) / 2
    total.backward()
    checked = 0
    for n, p in ours.named_parameters():
        if p.grad is None:
            continue
        g = g0[n].to(p.grad.device)
        den = max(p.grad.abs().max().item(), 1e-6)
        assert (g - p.grad).abs().max().item() / den <= 2e-3, n
        checked += 1
    assert checked > 40 and "position_relation_embedding.pos_proj.0.weight" in g0
