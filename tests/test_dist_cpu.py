"""CPU (-m "not gpu"): the data-parallel plumbing (unetb200.dist) with world_size 2 over gloo: bucket planning,
bucketed gradient SUM all-reduce in backward order, scalar loss SUM (UNet/model.py:233), ON_READ/MEAN moving statistics."""
import os
import socket
from collections import OrderedDict

import pytest
import torch
import torch.multiprocessing as mp


class _L:
    def __init__(self, name, b, e):
        self.name, self.seg_begin, self.seg_end = name, b, e


class _FakeModel:
    """just the attributes dist.py touches: layers (forward order, flat offsets descending), flat buffers"""

    def __init__(self, sizes):
        names = [f"l{i}" for i in range(len(sizes))]
        self.layers = OrderedDict()
        off = sum(sizes)
        for n, s in zip(names, sizes):          # forward order; the LAST layer sits at offset 0 (backward order layout)
            off -= s
            self.layers[n] = _L(n, off, off + s)
        self.n = sum(sizes)
        self.G = torch.zeros(self.n)
        self.P = torch.zeros(self.n)
        self.MM = torch.zeros(8)
        self.MV = torch.ones(8)
        self.changed = 0
        self._inference_stale = False

    def _weights_changed(self):
        self.changed += 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from unetb200 import dist as D
    dp = D.DataParallel(backend="gloo")
    try:
        sizes = [10, 3_000_000, 5, 1_500_000, 700_000, 40]
        m = _FakeModel(sizes)
        g = torch.Generator().manual_seed(100 + rank)
        m.G.copy_(torch.rand(m.n, generator=g))
        mine = m.G.clone()
        others = []
        for r in range(world):
            others.append(torch.rand(m.n, generator=torch.Generator().manual_seed(100 + r)))
        dp.begin_step(m)
        nb = len(dp._buckets)
        for name in reversed(list(m.layers)):                # backward order
            dp.layer_done(m, name)
        dp.finish_step(m)
        ok_sum = torch.allclose(m.G, sum(others))
        covered = sorted((lo, hi) for lo, hi, _ in dp._buckets)
        contiguous = covered[0][0] == 0 and covered[-1][1] == m.n and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
        loss = dp.reduce_sum(torch.tensor(float(rank + 1)))
        m.MM.fill_(float(rank))
        dp.average_moving_stats(m)
        m.P.fill_(float(rank + 5))
        dp.broadcast_params(m)
        q.put((rank, ok_sum, contiguous, nb, dp.allreduce_calls, float(loss), float(m.MM[0]), float(m.P[0]), m.changed,
               dp.num_replicas_in_sync, bool(torch.equal(mine, others[rank]))))
    finally:
        dp.shutdown()


def test_plan_buckets_contiguous_and_ordered():
    from unetb200.dist import plan_buckets
    layers = [("a", 0, 10), ("b", 10, 3_000_010), ("c", 3_000_010, 3_000_020), ("d", 3_000_020, 3_000_030)]
    b = plan_buckets(layers, min_params=1000)
    assert b == [(0, 3_000_010, "b"), (3_000_010, 3_000_030, "d")]
    assert plan_buckets(layers, min_params=1) == [(0, 10, "a"), (10, 3_000_010, "b"), (3_000_010, 3_000_020, "c"), (3_000_020, 3_000_030, "d")]


@pytest.mark.timeout(300)
def test_bucketed_allreduce_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_sum, contiguous, nb, calls, loss, mm, p0, changed, nrep, same in res:
        assert ok_sum, "bucketed all-reduce != sum of per-replica gradients"
        assert contiguous and nb >= 2 and calls == nb
        assert loss == 3.0                       # SUM of per-replica losses
        assert mm == 0.5                         # MEAN of moving statistics
        assert p0 == 5.0 and changed == 1        # rank 0's variables broadcast
        assert nrep == 2 and same
