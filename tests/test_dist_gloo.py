"""world_size-2 gloo test of the multi-GPU host logic (image sharding + MAX-over-ranks timing)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from relation_detr_b200 import dist as rdist
from relation_detr_b200 import workloads


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total_images, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, lr, w = rdist.init_process_group("gloo")
    assert (r, w) == (rank, world)
    lo, hi = rdist.shard_range(total_images, r, w)
    # every rank generates the same seeded global batch and keeps its own images: the union of the
    # shards must be the global batch, with no exchange on the data path
    boxes = workloads.make_boxes(total_images, 5, seed=3)
    mine = boxes[lo:hi]
    gathered = [None] * w
    dist.all_gather_object(gathered, (lo, hi, mine))
    rdist.barrier()
    slowest = rdist.max_over_ranks(10.0 + rank)
    total = rdist.sum_over_ranks(hi - lo)
    if rank == 0:
        whole = torch.cat([g[2] for g in gathered], 0)
        out_q.put((slowest, total, bool(torch.equal(whole, boxes)), [g[:2] for g in gathered]))
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing_reduction():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 7, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    slowest, total, same, spans = q.get()
    assert slowest == 11.0  # MAX over ranks, not rank 0's own time
    assert total == 7.0 and same
    assert spans == [(0, 4), (4, 7)]  # ragged split
