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
        with dp.moving_stats_averaged(m):
            mm_avg = float(m.MM[0])
        mm_after = float(m.MM[0])
        m.P.fill_(float(rank + 5))
        dp.broadcast_params(m)
        q.put((rank, ok_sum, contiguous, nb, dp.allreduce_calls, float(loss), (mm_avg, mm_after), float(m.P[0]), m.changed,
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
        assert mm == (0.5, float(rank))          # MEAN of moving statistics while read; each replica's own value afterwards
        assert p0 == 5.0 and changed == 1        # rank 0's variables broadcast
        assert nrep == 2 and same


def _banded_worker(rank, world, port, q):
    """segment_sharded over gloo with the kernels replaced by markers: the z-score sums carry the rank, the argmax kernel paints
    the zones it is handed with rank + 1 -- so the statistics all-reduce, the band geometry and the assembly on rank 0 are checked"""
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import numpy as np
    import unetb200._C as C
    import unetb200.inference as I
    from unetb200 import dist as D
    from unetb200 import model as M
    dp = D.DataParallel(backend="gloo")
    try:
        real = torch.cuda.is_available
        torch.cuda.is_available = lambda: True

        class Dry(M.UNet):
            def _stream(self):
                return 0

            def _call(self, name, *args):
                if name == "ub_head_argmax":
                    geo, n, mask, ld = args[9], args[6], args[10], args[11]
                    for cy0, cy1, cx0, cx1, dy, dx in geo[:n].tolist():
                        mask[dy:dy + cy1 - cy0, dx:dx + cx1 - cx0] = rank + 1
                return 0
        try:
            m = Dry(2, 1, 1, precision="bf16", seed=0, device="cpu")
        finally:
            torch.cuda.is_available = real
        seen = {}

        def fake_call(name, *args):
            if name == "ub_zscore_sums":
                args[2][:] = torch.tensor([[float(rank + 1), 10.0 * (rank + 1)]], dtype=torch.float64)
            elif name == "ub_zscore_apply_sums":
                seen["sums"] = args[3].clone()
                args[2].zero_()
            return 0
        C.call = fake_call
        H, W = 2048 + 150, 1024 + 490
        raw = torch.zeros((1, H, W), dtype=torch.int16)
        out = I.segment_sharded(raw, m, dp, 1024, radius=96, tile_batch=4)
        Hp, Wp = H + (16 - H % 16) % 16, W + (16 - W % 16) % 16
        shards = I.shard_plan(Hp, Wp, 1024, 96, world)
        want = np.zeros((Hp, Wp), dtype=np.uint8)
        for r, b in enumerate(shards):
            for t in b["tiles"]:
                want[t["dy"]:t["dy"] + t["cy1"] - t["cy0"], t["dx"]:t["dx"] + t["cx1"] - t["cx0"]] = r + 1
        ok_mask = bool(np.array_equal(out.numpy(), want[:H, :W]))          # the mask all-reduce leaves the assembled result on every rank
        q.put((rank, out is not None, ok_mask, seen["sums"].tolist()))
    finally:
        dp.shutdown()


@pytest.mark.timeout(300)
def test_banded_inference_world3_gloo():
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_banded_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, has_out, ok_mask, sums in res:
        assert has_out and ok_mask, "the all-reduced mask does not tile the image by owner"
        assert sums == [[6.0, 60.0]]             # 1 + 2 + 3 and 10 + 20 + 30: every rank normalises with the GLOBAL statistics
