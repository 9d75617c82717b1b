"""The N>1 host path on CPU: world_size-2 gloo process group exercising the batch sharding, the
rank-ordered gather and the max-over-ranks timing reduction that bench.py uses (no collective sits
on the data path of the kernels themselves)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    from wav2lip_b200.parallel import max_over_ranks, shard_range, sharded_forward, sum_over_ranks
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x = torch.arange(n * 3, dtype=torch.float32).reshape(n, 3)
        y = torch.arange(n * 2, dtype=torch.float32).reshape(n, 2) * 10

        def fake_forward(a, b):  # stands in for the per-rank kernel call: row-wise, no cross-sample op
            return torch.cat([a * 2 + 1, b.sum(dim=1, keepdim=True)], dim=1)

        full = sharded_forward(fake_forward, (x, y))
        ok = torch.equal(full, fake_forward(x, y))
        b, e = shard_range(n, rank, world)
        t = max_over_ranks(1.0 + rank)
        units = sum_over_ranks(float(e - b))
        dist.barrier()
        q.put((rank, ok, t, units))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [7, 128])
def test_world2_shard_gather_and_timing(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, t, units in res:
        assert ok, f"rank {rank}: gathered result differs from the unsharded forward"
        assert t == 2.0          # max over ranks of (1 + rank)
        assert units == float(n)  # every item processed exactly once
