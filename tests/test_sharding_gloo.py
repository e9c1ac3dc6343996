"""Host-side sharding logic of the multi-GPU path on CPU: two gloo ranks split a batch into contiguous slices,
"verify" their slice with a stand-in (the CPU cannot run the engine) and reassemble the status vector. CPU only."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    sys.path.insert(0, os.path.join(ROOT, "agora-blsful_b200"))
    import torch.distributed as dist
    import blsful_b200 as B
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    truth = (np.arange(n) % 7 == 3).astype(np.uint8)  # the "invalid" items of the whole batch

    def verify_slice(lo, hi):
        return truth[lo:hi]

    def gather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    full = B.verify_batch_sharded(verify_slice, n, rank, world, gather)
    lo, hi = B.shard_range(n, rank, world)
    q.put((rank, lo, hi, bool((full == truth).all())))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_covers_batch():
    sys.path.insert(0, os.path.join(ROOT, "agora-blsful_b200"))
    import blsful_b200 as B
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 4, 8):
            edges = [B.shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        B.shard_range(10, 2, 2)


@pytest.mark.timeout(120)
def test_two_rank_gloo_reassembles_status_vector():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n, world = 1001, 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == n
    assert all(r[3] for r in res)
