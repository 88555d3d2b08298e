"""CPU-only, world_size 2 over gloo: the gradient-bucket logic of the data-parallel path (one bucket per layer, in
backward completion order, summed over ranks, tiling the flat gradient vector)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multi_modal_transformers_tokenmerge_b200.parallel import GradBucketReducer, layer_buckets


def test_layer_buckets_tile_and_order():
    b = layer_buckets([100, 350, 600], 850)
    assert b == [(600, 850), (350, 600), (100, 350), (0, 100)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 850
        g = torch.arange(n, dtype=torch.float32) * (rank + 1)
        red = GradBucketReducer(g, layer_buckets([100, 350, 600], n))
        red.reduce()
        want = torch.arange(n, dtype=torch.float32) * sum(r + 1 for r in range(world))
        q.put((rank, bool(torch.equal(g, want))))
        with pytest.raises(AssertionError):
            GradBucketReducer(g, [(0, 10), (20, n)])
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
