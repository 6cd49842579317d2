"""Dry run of the training / test / inference launch schedules on the CPU: every `UNet._call` is intercepted, checked against
the C ABI declared in include/unetb200.h (entry point exists, argument count, pointer vs scalar kinds) and recorded -- no
kernel runs.  It guards the host-side wiring (a renamed entry point, a missing argument, a tensor passed where an int is
expected) for the default schedule and for the optional ones (side-stream weight gradients, fused reductions, BatchNorm fold)."""
import ctypes

import numpy as np
import pytest
import torch

import unetb200._C as C
from unetb200 import model as M


class DryUNet(M.UNet):
    def __init__(self, *a, **kw):
        self.calls = []
        self.args = []
        real = torch.cuda.is_available
        torch.cuda.is_available = lambda: True
        try:
            super().__init__(*a, device="cpu", **kw)
        finally:
            torch.cuda.is_available = real
        self.use_graph = False

    def _stream(self):
        return 0

    def _fork_side(self):
        return None                      # torch.cuda.stream(None) is a no-op context

    def _join_side(self):
        pass

    def _call(self, name, *args):
        assert name in C.DECLS, f"{name} is not declared in include/unetb200.h"
        params = C.DECLS[name][1]
        assert len(args) + 1 == len(params), f"{name}: {len(args)} arguments + stream, the header declares {len(params)}"
        for a, (ctype, pname) in zip(args, params):
            if ctype is ctypes.c_void_p:
                assert a is None or isinstance(a, int) or hasattr(a, "data_ptr"), f"{name}({pname}): pointer argument got {type(a).__name__}"
            elif ctype in (ctypes.c_float, ctypes.c_double):
                assert isinstance(a, (int, float)) and not isinstance(a, bool), f"{name}({pname}): float argument got {type(a).__name__}"
            else:
                assert isinstance(a, (int, np.integer)) and not isinstance(a, bool), f"{name}({pname}): integer argument got {type(a).__name__}"
        assert params[-1][0] is ctypes.c_void_p
        self.launches += 1
        self.calls.append((name, self._cur))
        self.args.append(args)
        return 0


def _step(m, N=1, C_=1, H=32, W=48, K=2):
    x = torch.zeros((N, C_, H, W), dtype=torch.float32)
    lab = torch.zeros((N, H, W), dtype=torch.uint8)
    m.calls.clear()
    m.args.clear()
    m.train_step(x, lab)
    return [n for n, _ in m.calls]


def _count(names):
    out = {}
    for n in names:
        out[n] = out.get(n, 0) + 1
    return out


@pytest.mark.parametrize("overlap", [False, True])
def test_unfolded_training_schedule(overlap):
    m = DryUNet(2, 1, 1, precision="bf16", seed=0)
    m.overlap_wgrad = overlap
    m.fold_bn = False                     # UB_FOLD_BN=0: every BatchNorm output is materialised
    c = _count(_step(m))
    # the tcgen05 forwards finalise their BatchNorm in the same launch (ub_*_fwd_bn); the first layer and the head keep ub_bn_finalize
    assert c["ub_conv3x3_fwd_bn"] == 17 and c["ub_conv_first_fwd"] == 1 and c["ub_deconv2x2_fwd_bn"] == 4 and c["ub_head_fwd"] == 1
    assert "ub_conv3x3_fwd" not in c and "ub_deconv2x2_fwd" not in c
    assert c["ub_conv3x3_wgrad"] == 17 and c["ub_deconv2x2_wgrad"] == 4 and c["ub_conv_first_wgrad"] == 1
    assert c.get("ub_conv3x3_dgrad", 0) + c.get("ub_conv3x3_dgrad_bnred", 0) == 17 and c["ub_deconv2x2_dgrad"] == 4
    assert c["ub_bn_apply"] + c["ub_bn_apply_pool"] == 22 and c["ub_bn_bwd_apply"] == 22 and c["ub_bn_finalize"] == 2
    assert c["ub_adam"] == 1 and c["ub_transpose_pack_multi"] == 1 and c["ub_maxpool2x2_bwd_add"] == 4
    assert "ub_fold_conv3_weights" not in c and "ub_maxpool2x2_bwd_add_bnred" not in c
    # the fp32 check mode and the class-weighted 3-channel, 8-class configuration walk the same graph
    for kw, shape in ((dict(precision="fp32"), (1, 1, 32, 32, 2)), (dict(precision="bf16", class_weights=[1.0] * 8), (1, 3, 32, 32, 8))):
        mm = DryUNet(shape[4], 1, shape[1], seed=0, **kw)
        mm.overlap_wgrad = overlap
        assert len(_step(mm, *shape)) > 150


def test_separate_finalize_schedule():
    m = DryUNet(2, 1, 1, precision="bf16", seed=0)
    m.fuse_finalize = False               # UB_FUSE_FINALIZE=0
    c = _count(_step(m))
    assert c["ub_conv3x3_fwd_cases"] == 13 and c["ub_conv3x3_fwd"] == 4 and c["ub_deconv2x2_fwd"] == 4 and c["ub_bn_finalize"] == 23
    assert "ub_conv3x3_fwd_bn" not in c


def test_optional_schedules():
    m = DryUNet(2, 1, 1, precision="bf16", seed=0)
    m.fuse_bn_reduce_ew = True
    m.fold_bn = False
    c = _count(_step(m))
    assert c["ub_maxpool2x2_bwd_add_bnred"] == 4 and c["ub_head_bwd_apply_bnred"] == 1 and "ub_maxpool2x2_bwd_add" not in c
    m = DryUNet(2, 1, 1, precision="bf16", seed=0)
    assert m.fold_bn and not m.bn_algebra # the default schedule folds BatchNorm into the consumer convolutions
    m.bn_algebra = True                   # UB_BN_ALGEBRA=1 (optional: BatchNorm-backward sums from the consumer's weight gradient)
    names = _step(m)
    c = _count(names)
    # 17 producers lose their BatchNorm-apply pass (13 conv/deconv layers, enc1b-3b behind a y-less pool, dec1b behind the folded
    # head); the 13 consumer convs fold the weights, use the case bias and fix the weight gradient
    assert c["ub_fold_conv3_weights"] == 13 and c["ub_conv3x3_fwd_bn"] == 17 and "ub_conv3x3_fwd_cases" not in c
    assert c["ub_bn_apply"] == 4 and c["ub_bn_apply_pool"] == 1 and c["ub_bn_pool"] == 3          # botb, dec2b-4b | enc4b | enc1b-3b
    assert c["ub_border_sums"] == 13 and c["ub_wgrad_fold_fix"] == 13 and c["ub_conv3x3_wgrad"] == 17
    assert c["ub_fold_head_weights"] == 1 and c["ub_head_wgrad_fold_fix"] == 1 and c["ub_head_fwd"] == 1
    # 14 BatchNorm layers get dbeta / dgamma from their consumer's weight gradient: enc<l>a, bota, dec<l>a, up<l> and dec1b (head);
    # the 8 others (encoder skips, layers in front of dropout or of a transposed convolution) keep a reduction over the gradient
    assert c["ub_bn_bwd_sums_wgrad"] == 14
    assert c.get("ub_bn_bwd_reduce", 0) + c.get("ub_conv3x3_dgrad_bnred", 0) == 8 and c["ub_bn_bwd_apply"] == 22
    before = [n for n, _ in m.calls]
    for i, n in enumerate(before):           # the sums are taken from dW_a, i.e. before the fold fix-up rewrites the weight gradient
        if n == "ub_bn_bwd_sums_wgrad" and m.args[i][5] == 9:
            assert "ub_wgrad_fold_fix" in before[i + 1:i + 3] and before[i - 1] in ("ub_border_sums", "ub_bn_bwd_sums_wgrad")
    m2 = DryUNet(2, 1, 1, precision="bf16", seed=0)          # the default
    c2 = _count(_step(m2))
    assert "ub_bn_bwd_sums_wgrad" not in c2 and c2.get("ub_bn_bwd_reduce", 0) + c2.get("ub_conv3x3_dgrad_bnred", 0) == 22
    # the 64 -> 64 dgrads (enc1b, dec1b) carry the sums of enc1a / dec1a by default
    assert c2["ub_conv3x3_dgrad_bnred"] == 12 and c2["ub_bn_bwd_reduce"] == 10 and c2["ub_deconv2x2_dgrad"] == 4
    m3 = DryUNet(2, 1, 1, precision="bf16", seed=0)
    m3.fuse_bn_reduce_deconv = True         # UB_FUSE_RED_DECONV=1: dec2b / dec3b / dec4b get theirs from the transposed-convolution dgrads
    assert m3.fuse_bn_reduce_64 == 2        # default: up1 gets its sums from the dgrad of dec1a where the row-streaming kernel takes it
    c3 = _count(_step(m3, H=16, W=128))
    assert c3["ub_deconv2x2_dgrad_bnred"] == 3 and c3["ub_deconv2x2_dgrad"] == 1 and c3["ub_conv3x3_dgrad_bnred"] == 13
    assert c3["ub_bn_bwd_reduce"] == 6      # the four encoder skips (pool backward), dec1b (head backward), botb (dropout backward)
    c3 = _count(_step(m3, H=32, W=48))      # a width the row-streaming kernel leaves to the pair kernel: dec1a's dgrad stays plain
    assert c3["ub_conv3x3_dgrad_bnred"] == 12 and c3["ub_bn_bwd_reduce"] == 7
    folded = {l for (n, l), a in zip(m.calls, m.args) if n == "ub_conv3x3_fwd_bn" and a[6] == 1}
    assert folded == {"enc1b", "enc2b", "enc3b", "enc4b", "botb", "dec4a", "dec4b", "dec3a", "dec3b", "dec2a", "dec2b", "dec1a", "dec1b"}
    # back to the default schedule on the same object
    m.fold_bn = False
    assert "ub_fold_conv3_weights" not in _count(_step(m)) and all(L.fold is None for L in m.layers.values())


def test_test_step_and_inference_schedules():
    m = DryUNet(2, 1, 1, precision="bf16", seed=0)
    x = torch.zeros((1, 1, 32, 32))
    m.calls.clear()
    m.test_step(x, torch.zeros((1, 32, 32), dtype=torch.uint8))
    c = _count(n for n, _ in m.calls)
    assert c["ub_conv3x3_fwd_affine"] == 17 and c["ub_maxpool2x2_fwd"] == 4 and "ub_bn_apply" not in c and c["ub_head_loss"] == 1
    m.calls.clear()
    m.forward_softmax(x)
    assert _count(n for n, _ in m.calls)["ub_head_loss"] == 1
    with pytest.raises(IOError):
        m.forward_softmax(torch.zeros((1, 1, 24, 32)))              # not a multiple of 16 (UNet/inference.py:42)
    with pytest.raises(IOError):
        m.train_step(torch.zeros((1, 2, 32, 32)), torch.zeros((1, 32, 32), dtype=torch.uint8))      # wrong channel count


def test_tiled_inference_schedule():
    """unetb200.inference.segment_device on the dry model: tiles grouped by shape, batched, every zone handed to ub_head_argmax
    exactly once with the geometry of tile_plan (UNet/inference.py:56-129); two ranks split the tiles without overlap"""
    import unetb200.inference as I

    class Geo(DryUNet):
        def _call(self, name, *args):
            if name == "ub_head_argmax":
                geo, n = args[9], args[6]
                self.zones.extend(geo[:n].tolist())
            return super()._call(name, *args)

    H, W = 2048 + 160, 1024 + 496
    plan = I.tile_plan(H, W, 1024, 96)
    seen = []
    for rank in (0, 1):
        m = Geo(2, 1, 1, precision="bf16", seed=0)
        m.zones = []

        class D:
            world_size = 2
        D.rank = rank
        real = torch.distributed.all_reduce
        torch.distributed.all_reduce = lambda *a, **k: None          # the mask SUM-reduce across ranks (no process group here)
        try:
            mask = I.segment_device(torch.zeros((1, H, W)), m, 1024, radius=96, tile_batch=4, dist=D)
        finally:
            torch.distributed.all_reduce = real
        assert tuple(mask.shape) == (H, W) and mask.dtype == torch.uint8
        c = _count(n for n, _ in m.calls)
        shapes = {}
        for t in plan[rank::2]:
            shapes.setdefault((t["y1"] - t["y0"], t["x1"] - t["x0"]), []).append(t)
        assert c["ub_head_argmax"] == sum(-(-len(v) // 4) for v in shapes.values())
        seen += m.zones
    want = sorted([t["cy0"], t["cy1"], t["cx0"], t["cx1"], t["dy"], t["dx"]] for t in plan)
    assert sorted(seen) == want
    cover = np.zeros((H, W), dtype=np.int32)
    for cy0, cy1, cx0, cx1, dy, dx in seen:
        cover[dy:dy + cy1 - cy0, dx:dx + cx1 - cx0] += 1
    assert (cover == 1).all()


def test_sharded_inference_schedule():
    """unetb200.inference.segment_sharded (a run of tiles per rank) on the dry model: per rank, the rows uploaded, the statistics call
    over the owned rows only, tiles read in place from the uploaded rows; the zones of all ranks tile the padded image"""
    import torch.distributed as td
    import unetb200.inference as I

    H, W, world = 2048 + 150, 1024 + 490, 3                      # not multiples of 16: bottom / right reflect padding
    Hp, Wp = H + (16 - H % 16) % 16, W + (16 - W % 16) % 16
    raw = torch.zeros((1, H, W), dtype=torch.int16)
    cover = np.zeros((Hp, Wp), dtype=np.int32)
    owned_rows = 0
    real_call, real_ar = C.call, td.all_reduce
    try:
        td.all_reduce = lambda *a, **k: None
        for rank in range(world):
            log = []

            def fake_call(name, *args):
                assert name in C.DECLS and len(args) == len(C.DECLS[name][1]), name          # called with the stream as last argument
                log.append((name, args))
                return 0

            C.call = fake_call

            class Geo(DryUNet):
                def _call(self, name, *args):
                    if name == "ub_head_argmax":
                        self.zones.extend(args[9][:args[6]].tolist())
                    if name == "ub_conv_first_fwd_affine_tiles":
                        self.origins.extend((args[1][:args[11]] + torch.tensor([self.row0, 0], dtype=torch.int32)).tolist())
                        self.extent = (args[2], args[3])
                    return super()._call(name, *args)

            m = Geo(2, 1, 1, precision="bf16", seed=0)
            m.zones, m.origins = [], []
            b = I.shard_plan(Hp, Wp, 1024, 96, world)[rank]
            m.row0 = b["y0"]

            class D:
                world_size = world
            D.rank = rank
            out = I.segment_sharded(raw, m, D, 1024, radius=96, tile_batch=4)
            assert tuple(out.shape) == (H, W)
            names = [n for n, _ in log]
            assert names == ["ub_zscore_sums", "ub_zscore_apply_sums"]
            _, a = log[0]
            rows_owned = min(b["sy1"], H) - b["sy0"]
            assert a[5] == rows_owned * W and a[6] == (min(b["y1"], H) - b["y0"]) * W and a[4] == 1
            owned_rows += rows_owned
            assert log[1][1][4] == float(H) * float(W)                                      # statistics of the whole unpadded image
            assert sorted(m.origins) == sorted([t["y0"], t["x0"]] for t in b["tiles"])        # tiles read in place, band-relative rows
            assert m.extent == (H - b["y0"], W)                                               # mirror past the UNPADDED extent
            for cy0, cy1, cx0, cx1, dy, dx in m.zones:
                cover[dy:dy + cy1 - cy0, dx:dx + cx1 - cx0] += 1
    finally:
        C.call, td.all_reduce = real_call, real_ar
    assert owned_rows == H and (cover == 1).all()


def test_bench_family_map_covers_the_default_schedule():
    """bench.py's roofline object divides algorithmic FLOPs by the CUDA-event time of the entry points it knows by NAME: every
    tensor-core entry point of the default step must be in its FAMILY map (a renamed entry point once left `roofline.achieved` empty)"""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    m = DryUNet(2, 1, 1, precision="bf16", seed=0)
    names = set(_step(m))
    tensor = {n for n in names if any(k in n for k in ("conv3x3_fwd", "conv3x3_dgrad", "conv3x3_wgrad", "deconv2x2_fwd", "deconv2x2_dgrad", "deconv2x2_wgrad"))}
    assert tensor and tensor <= set(bench.FAMILY), tensor - set(bench.FAMILY)
