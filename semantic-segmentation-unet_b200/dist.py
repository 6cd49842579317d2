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

BUCKET_MIN_PARAMS = int(os.environ.get("UB_BUCKET_MIN_PARAMS", "2000000"))
# measurement only (bench.py attribution runs): launch no gradient all-reduce at all, to time the step's compute at N ranks without its communication
_NO_REDUCE = os.environ.get("UB_DP_NOREDUCE", "0") == "1"


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
        if _NO_REDUCE:
            return
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
        if model.G.is_cuda and not _NO_REDUCE:
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

    def moving_stats_averaged(self, model):
        """ON_READ / MEAN aggregation of the BN moving statistics: a context in which model.MM / model.MV hold
        the mean over replicas (for a checkpoint or an evaluation); each replica's own statistics are put back on exit, as
        reading a tf ON_READ variable leaves the per-replica values untouched."""
        import contextlib

        @contextlib.contextmanager
        def ctx():
            if self.world_size == 1:
                yield
                return
            keep = (model.MM.clone(), model.MV.clone())
            for t in (model.MM, model.MV):
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                t.div_(self.world_size)
            model._inference_stale = True
            try:
                yield
            finally:
                model.MM.copy_(keep[0])
                model.MV.copy_(keep[1])
                model._inference_stale = True
        return ctx()

    def verify_step(self, model, images, labels):
        """Numerical self-check of the data-parallel step (UNet/model.py:223, :230-235) on live ranks: runs the public
        train_step (bucketed all-reduce overlapped with backward; CUDA-graph replay when enabled), then repeats the same step
        on each rank WITHOUT the all-reduce from the same state and with the same dropout masks, all-gathers those local
        gradients and compares their sum with the all-reduced gradient buffer; also checks that the updated parameters are
        bit-identical on every rank.  The model state is restored afterwards.  Returns a dict (same on all ranks)."""
        x = model._prep_images(images)
        N, _, H, W = x.shape
        lab = model._prep_labels(labels, N, H, W).clone()
        names = ("P", "M", "V", "MM", "MV")
        keep = {a: getattr(model, a).clone() for a in names}
        sc = model.step_count

        def restore():
            for a, t in keep.items():
                getattr(model, a).copy_(t)
            model.step_count = sc
            model._weights_changed()

        out = {"world_size": self.world_size}
        grads = {}
        for mode in ("graph", "eager"):
            use_graph = model.use_graph
            model.use_graph = use_graph and mode == "graph"
            reps = 3 if mode == "graph" else 1          # first call per shape runs eagerly, the second captures, the third replays
            for _ in range(reps):
                restore()
                model.train_step(x, lab)
            model.use_graph = use_graph
            torch.cuda.synchronize()
            grads[mode] = model.G.clone()
            p_after = model.P.clone()
            chk = p_after.view(torch.int32).to(torch.int64).sum().reshape(1)
            allc = [torch.empty_like(chk) for _ in range(self.world_size)]
            if self.world_size > 1:
                dist.all_gather(allc, chk)
            else:
                allc = [chk]
            out[f"params_identical_{mode}"] = bool(all(int(c) == int(allc[0]) for c in allc))
        # the same step without the collective
        restore()
        model.step_count = sc + 1
        dm = model._make_drop_masks(N, H, W)
        d, model.dist = model.dist, None
        try:
            model._step_body(x, lab, N, H, W, dm, False, False, False)
        finally:
            model.dist = d
        torch.cuda.synchronize()
        local = model.G.clone()
        parts = [torch.empty_like(local) for _ in range(self.world_size)]
        if self.world_size > 1:
            dist.all_gather(parts, local)
        else:
            parts = [local]
        total = torch.zeros_like(local, dtype=torch.float64)
        for t in parts:
            total += t.double()
        den = float(total.norm())
        for mode, g in grads.items():
            out[f"grad_rel_l2_{mode}"] = float((g.double() - total).norm()) / max(den, 1e-300)
            out[f"grad_max_abs_{mode}"] = float((g.double() - total).abs().max())
        out["grad_max_ref"] = float(total.abs().max())
        out["graph_vs_eager_bit_identical"] = bool(torch.equal(grads["graph"], grads["eager"]))
        out["local_grad_norms"] = [float(t.double().norm()) for t in parts]
        restore()
        out["ok"] = bool(out["grad_rel_l2_graph"] < 1e-5 and out["grad_rel_l2_eager"] < 1e-5 and out["params_identical_graph"]
                         and out["params_identical_eager"])
        return out

    def broadcast_params(self, model):
        """replicas start from rank 0's variables, as MirroredStrategy guarantees"""
        if self.world_size > 1:
            for t in (model.P, model.MM, model.MV):
                dist.broadcast(t, src=0)
            model._weights_changed()

    def barrier(self):
        if self.world_size > 1:
            dist.barrier()

    def shutdown(self, model=None, timeout_s=60.0):
        """tear the process group down in dependency order: the captured step graphs hold NCCL kernels of this communicator, so
        they go first; then the device is drained, the ranks meet, and the group is destroyed.  A watchdog bounds the destroy
        (returns False if it had to give up) so that a wedged peer cannot hang a GPU box."""
        if not dist.is_initialized():
            return True
        if model is not None:
            model._graphs.clear()
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        try:
            dist.barrier()
        except Exception:
            pass
        import threading
        t = threading.Thread(target=dist.destroy_process_group, daemon=True)
        t.start()
        t.join(timeout_s)
        return not t.is_alive()
