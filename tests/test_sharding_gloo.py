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


class _StandInEngine:
    """Stands in for the GPU engine in the CPU test of verify_batch_folded: a slice's "partial result" is the XOR-sum of
    a per-item tag (valid items contribute 0), the fold accepts iff the XOR of all partial results is 0, and a slice of a
    failed batch "bisects" to its own bad items - the same three-phase protocol, no curve arithmetic."""

    def __init__(self, truth):
        self.truth, self.pending = truth, None

    def miller_partial(self, impl, scheme, pk, sg, data, off, fmt):
        lo = int(pk[0]) | (int(pk[1]) << 8) | (int(pk[2]) << 16)          # the test encodes the slice start in its first key
        n = off.size - 1
        self.pending = (lo, n)
        acc = 0
        for i in range(lo, lo + n):
            if self.truth[i]:
                acc ^= (i + 1) * 2654435761 & 0xFFFFFFFF
        return acc.to_bytes(576, "big"), bytes(96)

    def final_exp_is_one(self, impl, gts, sums):
        acc = 0
        for g in gts:
            acc ^= int.from_bytes(g, "big")
        return acc == 0

    def partial_finish(self, n, ok):
        lo, m = self.pending
        assert m == n
        return np.zeros(n, dtype=np.uint8) if ok else self.truth[lo:lo + n].copy()


def _fold_worker(rank, world, port, n, bad, q):
    sys.path.insert(0, os.path.join(ROOT, "agora-blsful_b200"))
    import torch.distributed as dist
    import blsful_b200 as B
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    truth = np.zeros(n, dtype=np.uint8)
    truth[bad] = 1

    def all_gather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    pk = np.zeros(48 * n, dtype=np.uint8)
    for i in range(n):                       # item i's "key" carries its own index (see _StandInEngine.miller_partial)
        pk[48 * i], pk[48 * i + 1], pk[48 * i + 2] = i & 255, (i >> 8) & 255, (i >> 16) & 255
    sg = np.zeros(96 * n, dtype=np.uint8)
    off = np.arange(n + 1, dtype=np.uint64)
    lo, st = B.verify_batch_folded(_StandInEngine(truth), 2, 0, pk, sg, np.zeros(n, dtype=np.uint8), off, rank, world, all_gather)
    q.put((rank, lo, st.tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("bad", [[], [3, 700]])
def test_two_rank_gloo_folded_batch(bad):
    """verify_batch_folded over two gloo ranks: slices, exchange of the partial results, one verdict on every rank, finish."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n, world = 1001, 2
    procs = [ctx.Process(target=_fold_worker, args=(r, world, port, n, bad, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=150) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    full = res[0][2] + res[1][2]
    assert res[0][1] == 0 and res[1][1] == len(res[0][2]) and len(full) == n
    assert [i for i in range(n) if full[i]] == bad
