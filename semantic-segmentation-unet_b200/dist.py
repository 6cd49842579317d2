"""Data-parallel plumbing: the counterpart of tf.distribute.MirroredStrategy in UNet/train.py:57-63 and of the implicit
gradient all-reduce inside optimizer.apply_gradients (UNet/model.py:223).

One process per GPU (torchrun / torch.multiprocessing), torch.distributed is the plumbing (NCCL over NVLink/NVSwitch on
the B200 box, gloo in the CPU tests).  Semantics kept from the reference:
  * each replica normalises its loss by the GLOBAL batch, gradients are SUM-reduced;
  * BatchNorm uses per-replica batch statistics; moving statistics are averaged only when read (test / checkpoint);
  * the scalar loss is SUM-reduced (UNet/model.py:233).
Unlike the reference (one monolithic all-reduce after backward) gradients are reduced per BUCKET on a side stream as
soon as the bucket's last layer has finished its backward: the flat gradient buffer is laid out in backward order, so a
bucket is one contiguous slice of it.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

BUCKET_MIN_PARAMS = 2_000_000


def plan_buckets(layers, min_params=BUCKET_MIN_PARAMS):
    """layers: iterable of (name, seg_begin, seg_end) in BACKWARD order (contiguous, ascending offsets).
    Returns [(lo, hi, last_layer_name)]: a bucket is closed once it holds >= min_params elements."""
    buckets = []
    lo = None
    for name, b, e in layers:
        if lo is None:
            lo = b
        if e - lo >= min_params:
            buckets.append((lo, e, name))
            lo = None
    if lo is not None:
        buckets.append((lo, e, name))
    return buckets


class DataParallel:
    def __init__(self, backend=None, device=None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world_size = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        self.backend = backend
        if backend == "nccl":
            torch.cuda.set_device(self.local_rank if device is None else device)
        if self.world_size > 1 and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            kw = {}
            if backend == "nccl":
                kw["device_id"] = torch.device("cuda", torch.cuda.current_device())
            import datetime
            kw["timeout"] = datetime.timedelta(seconds=int(os.environ.get("UB_DIST_TIMEOUT_S", "180")))   # fail fast instead of hanging a GPU box
            dist.init_process_group(backend=backend, rank=self.rank, world_size=self.world_size, **kw)
        self._comm_stream = None
        self._buckets = None
        self._pending = []
        self._next = 0
        self.allreduce_calls = 0

    @property
    def num_replicas_in_sync(self):          # tf.distribute.Strategy attribute used at UNet/train.py:61
        return self.world_size

    # ---- gradient buckets -------------------------------------------------------------------------------------
    def _ensure_plan(self, model):
        if self._buckets is None:
            order = list(reversed(list(model.layers.values())))          # backward order == flat-buffer order
            self._buckets = plan_buckets([(L.name, L.seg_begin, L.seg_end) for L in order])
            if model.G.is_cuda:
                self._comm_stream = torch.cuda.Stream(device=model.G.device)

    def begin_step(self, model):
        self._ensure_plan(model)
        self._next = 0
        self._pending = []

    def layer_done(self, model, name):
        """called by the backward schedule right after `name`'s gradients were enqueued on the compute stream"""
        if self.world_size == 1 or self._next >= len(self._buckets):
            return
        lo, hi, last = self._buckets[self._next]
        if name != last:
            return
        self._next += 1
        self._reduce_slice(model.G[lo:hi], getattr(model, "_wstream", None))

    def _reduce_slice(self, t, side_stream=None):
        self.allreduce_calls += 1
        if t.is_cuda:
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(t.device))
            side_ready = None
            if side_stream is not None:          # weight gradients are produced on the model's side stream
                side_ready = torch.cuda.Event()
                side_ready.record(side_stream)
            with torch.cuda.stream(self._comm_stream):
                self._comm_stream.wait_event(ready)
                if side_ready is not None:
                    self._comm_stream.wait_event(side_ready)
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
        else:
            self._pending.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=True))

    def finish_step(self, model):
        """all buckets reduced before the optimizer reads G"""
        while self._next < len(self._buckets):          # safety: anything not yet launched
            lo, hi, _ = self._buckets[self._next]
            self._next += 1
            self._reduce_slice(model.G[lo:hi], getattr(model, "_wstream", None))
        if model.G.is_cuda:
            torch.cuda.current_stream(model.G.device).wait_stream(self._comm_stream)
        for w in self._pending:
            w.wait()
        self._pending = []

    # ---- scalars / statistics ---------------------------------------------------------------------------------
    def reduce_sum(self, t):
        if self.world_size > 1:
            t = t.clone()
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t

    def reduce_max(self, t):
        if self.world_size > 1:
            t = t.clone()
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t

    def average_moving_stats(self, model):
        """ON_READ / MEAN aggregation of the BN moving statistics"""
        if self.world_size > 1:
            for t in (model.MM, model.MV):
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                t.div_(self.world_size)
            model._inference_stale = True

    def broadcast_params(self, model):
        """replicas start from rank 0's variables, as MirroredStrategy guarantees"""
        if self.world_size > 1:
            for t in (model.P, model.MM, model.MV):
                dist.broadcast(t, src=0)
            model._weights_changed()

    def barrier(self):
        if self.world_size > 1:
            dist.barrier()

    def shutdown(self, timeout_s=30.0):
        """tear the process group down; bounded, because destroying a communicator whose kernels were captured in CUDA graphs
        has been seen to block (release the graphs first: UNet._graphs.clear())"""
        if dist.is_initialized():
            import threading
            t = threading.Thread(target=dist.destroy_process_group, daemon=True)
            t.start()
            t.join(timeout_s)
