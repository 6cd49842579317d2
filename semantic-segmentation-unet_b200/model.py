"""B200-native restatement of the reference `UNet` class (UNet/model.py:19-256) -- same constructor, constants and
public methods; every op of the graph is a hand-written sm_100a kernel reached through the C ABI (`_C.call`).

torch is only the allocation shell here: tensors own device memory, `data_ptr()` is what crosses the boundary.
There is no autograd tape: the backward schedule of the fixed reference graph is written out.

Layout decisions:
  activations NHWC, bf16 (precision='bf16', tcgen05 path) or fp32 (precision='fp32', CUDA-core check mode);
  trainables in ONE flat fp32 buffer (`P`), gradients in a flat twin (`G`), Adam moments `M`/`V`, bf16 shadow `S`;
  per layer the segments are [kernel | bias | beta | gamma], layers in REVERSE forward order so that gradient
  buckets for the data-parallel all-reduce are contiguous prefixes of `G` as backward proceeds.
"""
from __future__ import annotations

import math
import os
from collections import OrderedDict

import numpy as np
import torch

from . import _C
from ._C import call as _raw_call

BN_EPS = 1e-3          # Keras BatchNormalization defaults
BN_MOMENTUM = 0.99
ADAM_B1, ADAM_B2, ADAM_EPS = 0.9, 0.999, 1e-7   # Keras Adam defaults (App. A.6)


def _pad8(n):
    return (n + 7) // 8 * 8


class _Layer:
    __slots__ = ("name", "kind", "cin", "cout", "level", "ksize", "off_w", "off_b", "off_beta", "off_gamma", "n_w",
                 "off_stat", "seg_begin", "seg_end", "c0", "c1", "fold")


class UNet:
    _BASELINE_FEATURE_DEPTH = 64     # UNet/model.py:20
    _KERNEL_SIZE = 3
    _DECONV_KERNEL_SIZE = 2
    _POOLING_STRIDE = 2
    SIZE_FACTOR = 16                 # UNet/model.py:25
    RADIUS = 96                      # UNet/model.py:26

    def __init__(self, number_classes, global_batch_size, number_channels, learning_rate=3e-4, label_smoothing=0,
                 *, precision="bf16", device=None, seed=None, class_weights=None, dist=None):
        if not torch.cuda.is_available():
            raise RuntimeError("unetb200 requires a CUDA device (sm_100a); there is no CPU fallback")
        if not 0.0 <= float(label_smoothing) <= 1.0:
            raise ValueError("label_smoothing must be in [0, 1]")
        self.label_smoothing = float(label_smoothing)          # UNet/model.py:65, :77 (the reference's train.py leaves it at 0)
        if number_classes < 1 or number_classes > _C.MACROS["UB_MAX_CLASSES_ANY"]:
            # labels and masks are uint8 on this path, as in the reference's databases (UNet/build_lmdb.py:151 forces uint8 masks)
            raise ValueError(f"number_classes must be in [1, {_C.MACROS['UB_MAX_CLASSES_ANY']}]")
        if number_channels < 1 or number_channels > _C.MACROS["UB_MAX_CHANNELS"]:
            raise ValueError(f"number_channels must be in [1, {_C.MACROS['UB_MAX_CHANNELS']}]")
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.number_channels = int(number_channels)
        self.number_classes = int(number_classes)
        self.learning_rate = float(learning_rate)
        self.global_batch_size = int(global_batch_size)
        self.precision = precision
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.act_dtype = torch.bfloat16 if precision == "bf16" else torch.float32
        self.act_code = _C.UB_BF16 if precision == "bf16" else _C.UB_F32
        self.dist = dist
        self.step_count = 0
        self.launches = 0
        self.profile = None
        self.nvtx = os.environ.get("UB_NVTX", "0") == "1"
        self._cur = ""                    # layer the current launches belong to (profiling tag)
        self._buf = {}
        self._inference_stale = True
        self._radius_cache = None
        self._wstream = None
        self.overlap_wgrad = os.environ.get("UB_OVERLAP_WGRAD", "1") == "1"   # bf16 path: weight gradients on a side stream (see _side)
        self.use_graph = True              # replay the training step from a CUDA graph (one capture per input shape)
        self._graphs = {}
        self._lr_ring = None
        self.fuse_bn_reduce = True        # bf16 path: BatchNorm-backward sums in the producing dgrad's epilogue
        # ... also for the 64 -> 64 layers (enc1b -> enc1a, dec1b -> dec1a), whose dgrad is the row-streaming kernel (csrc/conv3_rows.cuh)
        # (on by default: 21.81 -> 21.59 ms per step, profiles/r02_ab_runs.md block P; UB_FUSE_RED64=0 restores the separate reduction pass)
        # ... and in the transposed-convolution dgrads, for dec2b / dec3b / dec4b (botb sits behind a dropout backward): UB_FUSE_RED_DECONV
        self.fuse_bn_reduce_deconv = os.environ.get("UB_FUSE_RED_DECONV", "0") == "1"
        # 2 (default): also the dgrad of dec1a (64 -> 64 + 64), which carries the sums of up1 -- where the image width lets the row-streaming
        # kernel take it (two launches, igemm_conv3.cu); elsewhere the 128-column pair kernel would fetch the `a` tile without look-ahead
        self.fuse_bn_reduce_64 = int(os.environ.get("UB_FUSE_RED64", "2"))
        # bf16 folded path, optional (UB_BN_ALGEBRA=1): dbeta / dgamma of a BatchNorm whose only consumer is a folded convolution come from that
        # convolution's weight gradient and border sums (ub_bn_bwd_sums_wgrad) instead of a reduction pass over the gradient tensor.  Parity
        # green and 4.3 GB less HBM traffic per step, but the BatchNorm backward of layer L then has to wait for the weight gradient of layer
        # L + 1, which takes the weight gradients off the side stream: measured 24.22 ms against 22.89 ms per step (profiles/r02_ab_runs.md).
        self.bn_algebra = os.environ.get("UB_BN_ALGEBRA", "0") == "1"
        self._sums_ready = set()
        self.fuse_finalize = os.environ.get("UB_FUSE_FINALIZE", "1") == "1"   # bf16 path: the forward's last CTA finalises the BatchNorm statistics
        self.fuse_bn_reduce_ew = os.environ.get("UB_FUSE_EW", "0") == "1"   # ... and in the pool-backward / head-backward passes
        # bf16 training forward with the producers' BatchNorm folded into the consumer convolutions (_forward_train_folded):
        # 17 of the 22 BatchNorm-apply passes disappear.  Parity green on B200 (fold_* cases), 23.91 -> 22.75 ms per config-2 step
        # (profiles/r02_fold_ab.md); UB_FOLD_BN=0 restores the y-materialising forward for A/B runs.
        self.fold_bn = os.environ.get("UB_FOLD_BN", "1") == "1"
        self._build_layout()
        self.class_weights = None
        if class_weights is not None:
            self.class_weights = torch.tensor(np.asarray(class_weights, dtype=np.float32), device=self.device)
        self._init_params(seed)

    # ------------------------------------------------------------------------------------------------ plumbing
    def _call(self, name, *args):
        self.launches += 1
        if self.nvtx:                      # UB_NVTX=1: one NVTX range per entry point, named <layer>:<entry> (nsys / ncu --nvtx)
            torch.cuda.nvtx.range_push(f"{self._cur}:{name}")
            try:
                return _raw_call(name, *args, self._stream())
            finally:
                torch.cuda.nvtx.range_pop()
        if self.profile is not None:       # bench.py: CUDA events around every entry point
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            rc = _raw_call(name, *args, self._stream())
            b.record()
            self.profile.append((name, self._cur, a, b))
            return rc
        return _raw_call(name, *args, self._stream())

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    # Weight-gradient kernels run on a side stream: wgrad(L) only needs dz_L and is not on the critical path
    # (dz_L -> dgrad(L) -> BN backward of the layer below -> ...).  It is enqueued AFTER dgrad(L) and waits for it (two
    # persistent tensor-core kernels cannot share an SM, whichever starts first would delay the other), so that its
    # tensor-core work runs beside the HBM-bound BatchNorm-backward passes of the layer below.
    def _side(self):
        if self._wstream is None:
            self._wstream = torch.cuda.Stream(device=self.device)
        return self._wstream

    def _fork_side(self):
        """side stream waits for everything enqueued so far on the current stream"""
        ws = self._side()
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        ws.wait_event(ev)
        return ws

    def _join_side(self):
        if self._wstream is not None:
            ev = torch.cuda.Event()
            ev.record(self._wstream)
            torch.cuda.current_stream(self.device).wait_event(ev)

    def _ensure(self, name, numel, dtype):
        t = self._buf.get(name)
        if t is None or t.numel() < numel or t.dtype != dtype:
            if t is not None and self._graphs:
                # captured step graphs hold the old pointer: drop them (they are re-captured on the next step of their shape)
                # before the old allocation can be handed out again
                torch.cuda.synchronize(self.device)
                self._graphs.clear()
            t = torch.empty(max(int(numel), 16), dtype=dtype, device=self.device)
            self._buf[name] = t
        return t

    # ------------------------------------------------------------------------------------------------ layout
    def _build_layout(self):
        b, nc, K = self._BASELINE_FEATURE_DEPTH, self.number_channels, self.number_classes
        fwd = [("enc1a", "first", nc, b, 1), ("enc1b", "conv", b, b, 1),
               ("enc2a", "conv", b, 2 * b, 2), ("enc2b", "conv", 2 * b, 2 * b, 2),
               ("enc3a", "conv", 2 * b, 4 * b, 3), ("enc3b", "conv", 4 * b, 4 * b, 3),
               ("enc4a", "conv", 4 * b, 8 * b, 4), ("enc4b", "conv", 8 * b, 8 * b, 4),
               ("bota", "conv", 8 * b, 16 * b, 5), ("botb", "conv", 16 * b, 16 * b, 5)]
        for lvl in (4, 3, 2, 1):
            c = b * (1 << (lvl - 1))
            fwd += [(f"up{lvl}", "deconv", 2 * c, c, lvl), (f"dec{lvl}a", "conv", 2 * c, c, lvl), (f"dec{lvl}b", "conv", c, c, lvl)]
        fwd.append(("head", "head", b, K, 1))
        self.layers = OrderedDict()
        for name, kind, cin, cout, level in fwd:
            L = _Layer()
            L.name, L.kind, L.cin, L.cout, L.level = name, kind, cin, cout, level
            L.ksize = {"first": 9, "conv": 9, "deconv": 4, "head": 1}[kind]
            L.n_w = cout * L.ksize * cin
            L.c0 = L.c1 = 0
            L.fold = None
            self.layers[name] = L
        off = 0
        soff = 0
        for name in reversed(list(self.layers)):       # head first: bucket order == backward order
            L = self.layers[name]
            off = (off + 127) // 128 * 128      # kernel segments start 256-byte (bf16) / 512-byte (fp32) aligned: TMA rows of the
            L.seg_begin = off                   # weight matrix then never straddle 128-byte lines
            L.off_w = off
            off += _pad8(L.n_w)
            L.off_b = off
            off += _pad8(L.cout)
            L.off_beta = off
            off += _pad8(L.cout)
            L.off_gamma = off
            off += _pad8(L.cout)
            L.seg_end = off
            L.off_stat = soff
            soff += _pad8(L.cout)
        self.n_flat = off
        self.n_stat = soff
        dev = self.device
        self.P = torch.zeros(off, dtype=torch.float32, device=dev)
        self.G = torch.zeros(off, dtype=torch.float32, device=dev)
        self.M = torch.zeros(off, dtype=torch.float32, device=dev)
        self.V = torch.zeros(off, dtype=torch.float32, device=dev)
        self.S = torch.zeros(off, dtype=torch.bfloat16, device=dev) if self.precision == "bf16" else None
        self.MM = torch.zeros(soff, dtype=torch.float32, device=dev)      # moving_mean
        self.MV = torch.ones(soff, dtype=torch.float32, device=dev)       # moving_variance
        self.MR = torch.ones(soff, dtype=torch.float32, device=dev)       # 1/sqrt(moving_var + eps) for inference
        self.FS = torch.ones(soff, dtype=torch.float32, device=dev)       # inference fold: gamma * MR
        self.FB = torch.zeros(soff, dtype=torch.float32, device=dev)      #                 beta - moving_mean * FS
        self.mean = torch.zeros(soff, dtype=torch.float32, device=dev)    # batch statistics of the last training fwd
        self.rstd = torch.ones(soff, dtype=torch.float32, device=dev)
        wt_dtype = torch.bfloat16 if self.precision == "bf16" else torch.float32
        self.WT = {n: torch.zeros(L.n_w, dtype=wt_dtype, device=dev) for n, L in self.layers.items() if L.kind in ("conv", "deconv")}
        self.metrics = torch.zeros(2, dtype=torch.float32, device=dev)    # [loss, accuracy] of the last step
        # partial rows of every reduction: [UB_STATS_ROWS][<= 2 * 2048 channels], or the head's {dW[K][64], db[K]} rows
        self.partial = torch.zeros(_C.UB_STATS_ROWS * max(2 * 2048, K * 64 + K), dtype=torch.float32, device=dev)
        self.partial_red = torch.zeros(_C.UB_STATS_ROWS * 2 * 1024, dtype=torch.float32, device=dev)   # BN-backward sums produced by a fused dgrad
        self._red_ready = None                                                                          # layer whose sums partial_red holds
        self.red = torch.zeros(4096, dtype=torch.float32, device=dev)
        self.fin_counter = torch.zeros(4, dtype=torch.int32, device=dev)           # ub_*_fwd_bn: CTAs done (left zero by every launch)

    def trainable_count(self):
        return sum(L.n_w + 3 * L.cout for L in self.layers.values())

    # views into the flat buffers ---------------------------------------------------------------------
    def _seg(self, flat, off, n):
        return flat[off:off + n]

    def _w(self, L, flat=None):
        return self._seg(self.P if flat is None else flat, L.off_w, L.n_w)

    def _wptr(self, L):
        """weights as the conv kernels consume them: bf16 shadow (tcgen05 path) or the fp32 master (check mode)"""
        if self.precision == "bf16" and L.kind in ("conv", "deconv"):
            return self.S[L.off_w:L.off_w + L.n_w]
        return self.P[L.off_w:L.off_w + L.n_w]

    # ------------------------------------------------------------------------------------------------ params
    def _init_params(self, seed):
        """Keras initial state: glorot_uniform kernels, zero bias, gamma 1, beta 0, moving 0 / 1."""
        g = torch.Generator(device="cpu")
        g.manual_seed(int(seed) if seed is not None else int(np.random.SeedSequence().entropy % (1 << 62)))
        host = torch.zeros(self.n_flat, dtype=torch.float32)
        for L in self.layers.values():
            k2 = {"first": 9, "conv": 9, "deconv": 4, "head": 1}[L.kind]
            lim = math.sqrt(6.0 / (k2 * L.cin + k2 * L.cout))
            host[L.off_w:L.off_w + L.n_w] = (torch.rand(L.n_w, generator=g) * 2 - 1) * lim
            host[L.off_gamma:L.off_gamma + L.cout] = 1.0
        self.P.copy_(host)
        self._weights_changed()

    def _weights_changed(self):
        """refresh the bf16 shadow and the dgrad packs after the fp32 master changed outside of Adam"""
        if self.S is not None:
            self._call("ub_cast_bf16", self.P, self.S, self.n_flat)
        self._repack_dgrad()
        self._inference_stale = True
        self._radius_cache = None

    def _repack_dgrad(self):
        """dgrad weight packs of every layer, one launch (21 tiny transposes otherwise)"""
        if getattr(self, "_pack_jobs", None) is None:
            import struct
            rec, begin = [], 0
            for n, L in self.layers.items():
                if L.kind not in ("conv", "deconv"):
                    continue
                T, flip, layout = (9, 1, 0) if L.kind == "conv" else (4, 0, 1)
                tc, tr = (L.cin + 31) // 32, (L.cout + 31) // 32
                rec.append(struct.pack("<QQiiiiiiii", self._w(L).data_ptr(), self.WT[n].data_ptr(), L.cout, T, L.cin, flip, layout, tc, tr, begin))
                begin += tc * tr * T
            blob = np.frombuffer(b"".join(rec), dtype=np.uint8).copy()
            self._pack_jobs = torch.from_numpy(blob).to(self.device)
            self._pack_njobs, self._pack_tiles = len(rec), begin
        self._call("ub_transpose_pack_multi", self._pack_jobs, self._pack_njobs, self._pack_tiles, self.act_code)

    def _import_flat(self, params, host):
        """dict in the TF layouts (conv [kh,kw,Cin,Cout], deconv [kh,kw,Cout,Cin], head [1,1,Cin,K]) -> flat host buffer"""
        for n, L in self.layers.items():
            w = torch.as_tensor(np.asarray(params[n + "/kernel"], dtype=np.float32))
            want = (2, 2, L.cout, L.cin) if L.kind == "deconv" else ((1, 1) if L.kind == "head" else (3, 3)) + (L.cin, L.cout)
            if tuple(w.shape) != want:
                raise IOError(f"{n}/kernel has shape {tuple(w.shape)}, expected {want}")
            if L.kind == "deconv":
                packed = w.reshape(4 * L.cout, L.cin)
            else:
                packed = w.permute(3, 0, 1, 2).reshape(L.cout, -1)
            host[L.off_w:L.off_w + L.n_w] = packed.reshape(-1)
            host[L.off_b:L.off_b + L.cout] = torch.as_tensor(np.asarray(params[n + "/bias"], dtype=np.float32))
            host[L.off_beta:L.off_beta + L.cout] = torch.as_tensor(np.asarray(params[n + "/beta"], dtype=np.float32))
            host[L.off_gamma:L.off_gamma + L.cout] = torch.as_tensor(np.asarray(params[n + "/gamma"], dtype=np.float32))
        return host

    def load_oracle_params(self, params):
        """params: dict in the TF layouts used by the oracle / a Keras checkpoint (conv [kh,kw,Cin,Cout], deconv
        [kh,kw,Cout,Cin], head [1,1,Cin,K]); also moving_mean / moving_var."""
        host = self._import_flat(params, self.P.cpu())
        mm, mv = self.MM.cpu(), self.MV.cpu()
        for n, L in self.layers.items():
            if n + "/moving_mean" in params:
                mm[L.off_stat:L.off_stat + L.cout] = torch.as_tensor(np.asarray(params[n + "/moving_mean"], dtype=np.float32))
                mv[L.off_stat:L.off_stat + L.cout] = torch.as_tensor(np.asarray(params[n + "/moving_var"], dtype=np.float32))
        self.P.copy_(host)
        self.MM.copy_(mm)
        self.MV.copy_(mv)
        self._weights_changed()

    def _unpack(self, flat_host, L):
        w = flat_host[L.off_w:L.off_w + L.n_w]
        if L.kind == "deconv":
            return w.reshape(2, 2, L.cout, L.cin)
        k = 3 if L.kind in ("conv", "first") else 1
        return w.reshape(L.cout, k, k, L.cin).permute(1, 2, 3, 0).contiguous()

    def export_flat(self, flat, with_stats=False):
        """flat buffer (P or G) -> dict in TF layouts (for parity checks and checkpoints)"""
        host = flat.detach().float().cpu()
        out = OrderedDict()
        for n, L in self.layers.items():
            out[n + "/kernel"] = self._unpack(host, L).numpy()
            out[n + "/bias"] = host[L.off_b:L.off_b + L.cout].numpy().copy()
            out[n + "/gamma"] = host[L.off_gamma:L.off_gamma + L.cout].numpy().copy()
            out[n + "/beta"] = host[L.off_beta:L.off_beta + L.cout].numpy().copy()
        if with_stats:
            mm, mv = self.MM.cpu(), self.MV.cpu()
            for n, L in self.layers.items():
                out[n + "/moving_mean"] = mm[L.off_stat:L.off_stat + L.cout].numpy().copy()
                out[n + "/moving_var"] = mv[L.off_stat:L.off_stat + L.cout].numpy().copy()
        return out

    def export_params(self):
        return self.export_flat(self.P, with_stats=True)

    def export_grads(self):
        return self.export_flat(self.G)

    def export_activation_pattern(self, N, H, W):
        """({layer: bool NCHW [a > 0]}, {"pool<l>": uint8 NCHW window slot 2*dy+dx}) of the last training forward.
        Parity tests hand these to the oracle so that both sides differentiate the same piecewise-linear branch."""
        relu, pool = {}, {}
        for n, L in self.layers.items():
            if L.kind == "deconv":
                continue
            h, w = self._dims(H, W, L.level)
            a = self._b("a:" + n)[:N * h * w * L.cout].view(N, h, w, L.cout)
            relu[n] = (a > 0).permute(0, 3, 1, 2).contiguous().cpu()
        for lvl in (1, 2, 3, 4):
            h, w = self._dims(H, W, lvl + 1)
            C = self._BASELINE_FEATURE_DEPTH << (lvl - 1)
            idx = self._b(f"idx{lvl}")[:N * h * w * C].view(N, h, w, C)
            pool[f"pool{lvl}"] = idx.permute(0, 3, 1, 2).contiguous().cpu()
        return relu, pool

    # reference API ------------------------------------------------------------------------------------
    def get_optimizer(self):
        return self

    def set_learning_rate(self, learning_rate):          # UNet/model.py:154-155
        self.learning_rate = float(learning_rate)

    def get_learning_rate(self):
        return self.learning_rate

    def get_keras_model(self):                           # UNet/model.py:148-149
        return self._model_call

    @staticmethod
    def _round_radius(x):
        return int(UNet.SIZE_FACTOR * math.ceil(float(x) / UNet.SIZE_FACTOR))

    def input_gradient(self, images, dlogits_fn):
        """dL/d(images) of the inference-mode graph (training=False: moving statistics, no dropout) for a loss whose
        gradient w.r.t. the softmax INPUT (the head's BatchNorm output) is dlogits_fn(softmax [N,H,W,K] numpy) -> numpy.
        Runs on a private fp32 (check-mode) copy of this model: the reference differentiates in fp32 and thresholds the
        result at 1e-8 (UNet/model.py:165-202)."""
        m = UNet(self.number_classes, 1, self.number_channels, self.learning_rate, self.label_smoothing, precision="fp32", device=self.device, seed=0)
        m.P.copy_(self.P)
        m.MM.copy_(self.MM)
        m.MV.copy_(self.MV)
        m._weights_changed()
        x = m._prep_images(images)
        N, _, H, W = x.shape
        sm = m.forward_softmax(x, training=False)
        dl = torch.as_tensor(np.ascontiguousarray(dlogits_fn(sm.cpu().numpy()), dtype=np.float32)).to(m.device)
        m._ensure("dlogits", N * H * W * m.number_classes, torch.float32)[:dl.numel()].copy_(dl.reshape(-1))
        m._alloc(N, H, W, True)            # gradient buffers (forward_softmax allocated the inference set only)
        m._bwd_inference = True
        dx = torch.zeros_like(x)
        m._backward(x, N, H, W, None, input_grad=dx)
        return dx

    def estimate_radius(self, noise=None):
        """UNet/model.py:160-202: empirical receptive field -- gradient of a mean-absolute-error loss that is non-zero only
        at the centre pixel of a 192 x 192 noise image w.r.t. that image; radius = half the extent where |grad| > 1e-8,
        rounded up to a multiple of 16 (falls back to RADIUS = 96).  The reference draws unseeded noise and runs ten
        redundant forward passes (only the last tape is used); one pass is run here, `noise` may be injected for tests."""
        Nn = 2 * UNet.RADIUS
        if noise is None:
            if self._radius_cache is not None:      # same weights -> same receptive field: estimated once per weight state
                return self._radius_cache           # (the reference re-estimates it for every image, inference.py:54)
            noise = np.random.normal(size=(1, self.number_channels, Nn, Nn))
            cache = True
        else:
            cache = False
        mid = int(Nn / 2)
        K = self.number_classes

        def dlogits(sm):
            # loss = sum_pixels mean_k |msk - sm|, msk = sm except 1 - sm at the centre: only the centre pixel contributes
            p = sm[0, mid, mid].astype(np.float64)
            dsm = -np.sign((1.0 - p) - p) / K                      # d loss / d softmax_k at the centre
            dy = p * (dsm - float((dsm * p).sum()))                # through the softmax Jacobian
            out = np.zeros(sm.shape, dtype=np.float32)
            out[0, mid, mid] = dy
            return out

        g = self.input_gradient(np.asarray(noise, dtype=np.float32), dlogits)
        grad_img = np.abs(g[0].cpu().numpy())
        grad_img = np.average(grad_img, axis=0) if self.number_channels > 1 else grad_img.squeeze()
        print('Theoretical RF: {}'.format(UNet.RADIUS))
        vec = np.maximum(np.max(grad_img, axis=0).squeeze(), np.max(grad_img, axis=1).squeeze())
        idx = np.nonzero(vec > 1e-8)[0]
        if len(idx) < 2:
            radius = UNet.RADIUS
            print('ERF based radius detection failed, defaulting to theoretical radius: {}'.format(radius))
        else:
            erf = int((np.max(idx) - np.min(idx)) / 2)
            radius = UNet._round_radius(erf)
            print('computed radius : "{}"'.format(radius))
        self._last_erf_grad = grad_img
        if cache:
            self._radius_cache = radius
        return radius

    # ------------------------------------------------------------------------------------------------ shapes
    def _dims(self, H, W, level):
        s = 1 << (level - 1)
        return H // s, W // s

    def _alloc(self, N, H, W, training):
        adt = self.act_dtype
        for n, L in self.layers.items():
            if L.kind == "head":
                continue
            h, w = self._dims(H, W, L.level)
            numel = N * h * w * L.cout
            if training or self.precision != "bf16":
                self._ensure("a:" + n, numel, adt)      # bf16 inference folds BN into the producer: only y is kept
            self._ensure("y:" + n, numel, adt)
            if training:
                self._ensure("g:" + n, numel, adt)
        for lvl in (1, 2, 3, 4):
            h, w = self._dims(H, W, lvl + 1)
            C = self._BASELINE_FEATURE_DEPTH << (lvl - 1)
            self._ensure(f"pool{lvl}", N * h * w * C, adt)
            if training or self.precision != "bf16":
                self._ensure(f"idx{lvl}", N * h * w * C, torch.uint8)
            if training:
                self._ensure(f"gpool{lvl}", N * h * w * C, adt)
                self._ensure(f"gskip{lvl}", N * 4 * h * w * C, adt)
        P = N * H * W
        K = self.number_classes
        self._ensure("a:head", P * K, torch.float32)
        if training:
            self._ensure("dlogits", P * K, torch.float32)
            need = 0
            for n, L in self.layers.items():
                h, w = self._dims(H, W, L.level)
                if L.kind == "conv":
                    c0 = L.cin // 2 if n.startswith("dec") and n.endswith("a") else L.cin
                    need = max(need, _C.lib.ub_conv3x3_wgrad_workspace_bytes(c0, L.cin - c0, L.cout, N, h, w)
                               if self.precision == "bf16" else 0)
                elif L.kind == "deconv":
                    need = max(need, _C.lib.ub_deconv2x2_wgrad_workspace_bytes(L.cin, L.cout, N, h // 2, w // 2)
                               if self.precision == "bf16" else 0)
            self._ensure("wgrad_ws", max(need, 16), torch.uint8)
            self._ensure("first_ws", _C.UB_STATS_ROWS * self.number_channels * 9 * 64, torch.float32)

    def _b(self, name):
        return self._buf[name]

    # ------------------------------------------------------------------------------------------------ forward
    def _bn_vectors(self, L, training):
        o, c = L.off_stat, L.cout
        if training:
            return self.mean[o:o + c], self.rstd[o:o + c]
        return self.MM[o:o + c], self.MR[o:o + c]

    def _affine(self, L):
        return self.P[L.off_gamma:L.off_gamma + L.cout], self.P[L.off_beta:L.off_beta + L.cout]

    def _finalize(self, L, ncols, groups, count):
        o, c = L.off_stat, L.cout
        self._call("ub_bn_finalize", self.partial, ncols, groups, count, self.mean[o:o + c], self.rstd[o:o + c],
                   self.MM[o:o + c], self.MV[o:o + c], BN_MOMENTUM, BN_EPS)

    def _conv_fwd(self, L, x0, c0, x1, c1, N, h, w, training):
        """conv3x3 + bias + relu -> a:<name>; batch statistics -> mean/rstd (training)"""
        self._cur = L.name
        a = self._b("a:" + L.name)
        bias = self.P[L.off_b:L.off_b + L.cout]
        L.c0, L.c1 = c0, c1
        if self.precision == "bf16" and training and self.fuse_finalize:
            self._conv_fwd_bn(L, x0, c0, x1, c1, self._wptr(L), bias, 0, a, N, h, w)
            return a
        if self.precision == "bf16":
            self._call("ub_conv3x3_fwd", x0, c0, x1, c1, self._wptr(L), bias, a, self.partial if training else None,
                       N, h, w, L.cout, 1)
        else:
            self._call("ub_check_conv3x3", x0, c0, x1, c1, self._wptr(L), bias, a, L.cout, None, 0, N, h, w, 1)
            if training:
                self._call("ub_bn_stats", a, self.partial, N * h * w, L.cout, self.act_code)
        if training:
            self._finalize(L, L.cout, 1, N * h * w)
        return a

    def _conv_fwd_bn(self, L, x0, c0, x1, c1, wt, bias, bias_cases, a, N, h, w):
        """conv3x3 + bias + relu and, in the same launch, the batch statistics -> mean / rstd / moving statistics of the BatchNorm that
        follows (the last CTA finalises: no ub_bn_finalize launch, no zero-fill of partial rows)"""
        o, c = L.off_stat, L.cout
        self._call("ub_conv3x3_fwd_bn", x0, c0, x1, c1, wt, bias, bias_cases, a, self.partial, N, h, w, L.cout, 1,
                   self.mean[o:o + c], self.rstd[o:o + c], self.MM[o:o + c], self.MV[o:o + c], BN_MOMENTUM, BN_EPS, self.fin_counter)

    def _bn_apply(self, L, N, h, w, training, drop=None, pool_lvl=None):
        self._cur = L.name
        a, y = self._b("a:" + L.name), self._b("y:" + L.name)
        mean, rstd = self._bn_vectors(L, training)
        gamma, beta = self._affine(L)
        if pool_lvl is None:
            self._call("ub_bn_apply", a, y, mean, rstd, gamma, beta, drop, N * h * w, L.cout, self.act_code)
        else:
            self._call("ub_bn_apply_pool", a, y, self._b(f"pool{pool_lvl}"), self._b(f"idx{pool_lvl}"), mean, rstd, gamma, beta,
                       drop, N, h, w, L.cout, self.act_code)
        return y

    def _deconv_fwd_bn(self, Lu, x, z, N, hi, wi):
        o, c = Lu.off_stat, Lu.cout
        self._call("ub_deconv2x2_fwd_bn", x, Lu.cin, self._wptr(Lu), self.P[Lu.off_b:Lu.off_b + c], z, self.partial, N, hi, wi, c,
                   self.mean[o:o + c], self.rstd[o:o + c], self.MM[o:o + c], self.MV[o:o + c], BN_MOMENTUM, BN_EPS, self.fin_counter)

    def _forward(self, x, N, H, W, training, drop_masks=None, view=None):
        """x: fp32 NCHW device tensor, contiguous.  Leaves y:dec1b ready for the head; returns nothing.
        view (bf16 inference only, x = None): (img [C, rows, pitch] fp32, img_h, img_w, origins int32 [N, 2]) -- the N tiles are read
        in place from the resident normalised image, mirrored past (img_h, img_w) (ub_conv_first_fwd_affine_tiles)."""
        if H % self.SIZE_FACTOR or W % self.SIZE_FACTOR:
            raise IOError(f"Input image tile size needs to be a multiple of {self.SIZE_FACTOR} to allow integer sized downscaled feature maps")
        self._alloc(N, H, W, training)
        if not training and self._inference_stale:
            self._call("ub_bn_inference_rstd", self.MV, self.MR, self.n_stat, BN_EPS)
            for L in self.layers.values():
                o, c = L.off_stat, L.cout
                gamma, beta = self._affine(L)
                self._call("ub_bn_fold", gamma, beta, self.MM[o:o + c], self.MV[o:o + c], self.FS[o:o + c], self.FB[o:o + c], c, BN_EPS)
            self._inference_stale = False
        if not training and self.precision == "bf16":
            return self._forward_folded(x, N, H, W, view)
        assert view is None, "tile views are an inference feature of the bf16 path"
        if training and self.fold_bn and self.precision == "bf16":
            return self._forward_train_folded(x, N, H, W, drop_masks or {})
        for L in self.layers.values():
            L.fold = None
        Ls = self.layers
        dm = drop_masks or {}
        # ---- encoder
        L = Ls["enc1a"]
        self._cur = "enc1a"
        self._call("ub_conv_first_fwd", x, self.P[L.off_w:L.off_w + L.n_w], self.P[L.off_b:L.off_b + L.cout], self._b("a:enc1a"),
                   self.partial if training else None, N, H, W, self.number_channels, self.act_code)
        if training:
            self._finalize(L, L.cout, 1, N * H * W)
        cur = self._bn_apply(L, N, H, W, training)
        for lvl in (1, 2, 3, 4):
            h, w = self._dims(H, W, lvl)
            if lvl > 1:
                La = Ls[f"enc{lvl}a"]
                self._conv_fwd(La, self._b(f"pool{lvl - 1}"), La.cin, None, 0, N, h, w, training)
                cur = self._bn_apply(La, N, h, w, training)
            Lb = Ls[f"enc{lvl}b"]
            self._conv_fwd(Lb, cur, Lb.cin, None, 0, N, h, w, training)
            self._bn_apply(Lb, N, h, w, training, drop=dm.get("drop4") if (lvl == 4 and training) else None, pool_lvl=lvl)
        # ---- bottleneck
        h, w = self._dims(H, W, 5)
        La, Lb = Ls["bota"], Ls["botb"]
        self._conv_fwd(La, self._b("pool4"), La.cin, None, 0, N, h, w, training)
        cur = self._bn_apply(La, N, h, w, training)
        self._conv_fwd(Lb, cur, Lb.cin, None, 0, N, h, w, training)
        cur = self._bn_apply(Lb, N, h, w, training, drop=dm.get("dropb") if training else None)
        # ---- decoder
        for lvl in (4, 3, 2, 1):
            hi, wi = self._dims(H, W, lvl + 1)
            h, w = self._dims(H, W, lvl)
            Lu = Ls[f"up{lvl}"]
            self._cur = Lu.name
            z = self._b("a:" + Lu.name)
            bias = self.P[Lu.off_b:Lu.off_b + Lu.cout]
            if self.precision == "bf16" and training and self.fuse_finalize:
                self._deconv_fwd_bn(Lu, cur, z, N, hi, wi)
            elif self.precision == "bf16":
                self._call("ub_deconv2x2_fwd", cur, Lu.cin, self._wptr(Lu), bias, z, self.partial if training else None,
                           N, hi, wi, Lu.cout)
                if training:
                    self._finalize(Lu, 4 * Lu.cout, 4, N * h * w)
            else:
                self._call("ub_check_deconv2x2_fwd", cur, self._wptr(Lu), bias, z, N, hi, wi, Lu.cin, Lu.cout)
                if training:
                    self._call("ub_bn_stats", z, self.partial, N * h * w, Lu.cout, self.act_code)
                    self._finalize(Lu, Lu.cout, 1, N * h * w)
            u = self._bn_apply(Lu, N, h, w, training)
            La, Lb = Ls[f"dec{lvl}a"], Ls[f"dec{lvl}b"]
            skip = self._b(f"y:enc{lvl}b")
            self._conv_fwd(La, skip, Lu.cout, u, Lu.cout, N, h, w, training)       # concat [skip, up] (model.py:117)
            cur = self._bn_apply(La, N, h, w, training)
            self._conv_fwd(Lb, cur, Lb.cin, None, 0, N, h, w, training)
            cur = self._bn_apply(Lb, N, h, w, training)
        return cur

    def _conv_fwd_fold(self, L, x0, c0, p0, x1, c1, p1, N, h, w):
        """conv3x3 whose sources are pre-BatchNorm activations: p0 / p1 = the producer layer whose BatchNorm (batch statistics of
        this step) is folded into this layer's weights, or None for a source that is already normalised"""
        self._cur = L.name
        a = self._b("a:" + L.name)
        L.c0, L.c1, L.fold = c0, c1, (p0, p1)
        wf = self._ensure("wfold:" + L.name, L.n_w, torch.bfloat16)
        b9 = self._ensure("bias9:" + L.name, 9 * L.cout, torch.float32)
        sc = self._ensure("fold_s:" + L.name, L.cin, torch.float32)
        sh = self._ensure("fold_t:" + L.name, L.cin, torch.float32)

        def vecs(p):
            if p is None:
                return None, None, None, None
            mean, rstd = self._bn_vectors(p, True)
            gamma, beta = self._affine(p)
            return mean, rstd, gamma, beta

        self._call("ub_fold_conv3_weights", self._w(L), L.cout, c0, *vecs(p0), c1, *vecs(p1), self.P[L.off_b:L.off_b + L.cout], wf, b9, sc, sh)
        if self.fuse_finalize:
            self._conv_fwd_bn(L, x0, c0, x1, c1, wf, b9, 1, a, N, h, w)
            return a
        self._call("ub_conv3x3_fwd_cases", x0, c0, x1, c1, wf, b9, a, self.partial, N, h, w, L.cout, 1)
        self._finalize(L, L.cout, 1, N * h * w)
        return a

    def _forward_train_folded(self, x, N, H, W, dm):
        """Training forward (bf16) in which a layer's BatchNorm output is not written when all of its consumers are 3x3
        convolutions: those read the pre-BatchNorm activation with folded weights and a 9-case border bias (csrc/fold.cu).
        Folded producers: enc<l>a, bota, dec<l>a, up<l>, enc1b-enc3b (pool output written, y not) and dec1b (folded into the 1x1
        head).  Kept: layers followed by dropout (enc4b, botb) or by a transposed convolution (dec2b-dec4b)."""
        Ls = self.layers
        for L in Ls.values():
            L.fold = None
        L = Ls["enc1a"]
        self._cur = "enc1a"
        self._call("ub_conv_first_fwd", x, self.P[L.off_w:L.off_w + L.n_w], self.P[L.off_b:L.off_b + L.cout], self._b("a:enc1a"),
                   self.partial, N, H, W, self.number_channels, self.act_code)
        self._finalize(L, L.cout, 1, N * H * W)
        for lvl in (1, 2, 3, 4):
            h, w = self._dims(H, W, lvl)
            La, Lb = Ls[f"enc{lvl}a"], Ls[f"enc{lvl}b"]
            if lvl > 1:
                self._conv_fwd(La, self._b(f"pool{lvl - 1}"), La.cin, None, 0, N, h, w, True)
            self._conv_fwd_fold(Lb, self._b("a:" + La.name), Lb.cin, La, None, 0, None, N, h, w)
            if lvl == 4:          # dropout sits between this BatchNorm and its consumers: y is materialised
                self._bn_apply(Lb, N, h, w, True, drop=dm.get("drop4"), pool_lvl=lvl)
            else:                 # pool and skip only: the pooled tensor is written, the skip consumer folds this BatchNorm
                self._cur = Lb.name
                mean, rstd = self._bn_vectors(Lb, True)
                gamma, beta = self._affine(Lb)
                self._call("ub_bn_pool", self._b("a:" + Lb.name), self._b(f"pool{lvl}"), self._b(f"idx{lvl}"), mean, rstd, gamma, beta, None,
                           N, h, w, Lb.cout, self.act_code)
        h, w = self._dims(H, W, 5)
        La, Lb = Ls["bota"], Ls["botb"]
        self._conv_fwd(La, self._b("pool4"), La.cin, None, 0, N, h, w, True)
        self._conv_fwd_fold(Lb, self._b("a:bota"), Lb.cin, La, None, 0, None, N, h, w)
        cur = self._bn_apply(Lb, N, h, w, True, drop=dm.get("dropb"))
        for lvl in (4, 3, 2, 1):
            hi, wi = self._dims(H, W, lvl + 1)
            h, w = self._dims(H, W, lvl)
            Lu, La, Lb = Ls[f"up{lvl}"], Ls[f"dec{lvl}a"], Ls[f"dec{lvl}b"]
            self._cur = Lu.name
            if self.fuse_finalize:
                self._deconv_fwd_bn(Lu, cur, self._b("a:" + Lu.name), N, hi, wi)
            else:
                self._call("ub_deconv2x2_fwd", cur, Lu.cin, self._wptr(Lu), self.P[Lu.off_b:Lu.off_b + Lu.cout], self._b("a:" + Lu.name), self.partial,
                           N, hi, wi, Lu.cout)
                self._finalize(Lu, 4 * Lu.cout, 4, N * h * w)
            # concat [skip, up] (model.py:117): both BatchNorms are folded, except the level-4 skip (dropout: a real y tensor)
            Ls_skip = Ls[f"enc{lvl}b"]
            if lvl == 4:
                self._conv_fwd_fold(La, self._b("y:enc4b"), Lu.cout, None, self._b("a:" + Lu.name), Lu.cout, Lu, N, h, w)
            else:
                self._conv_fwd_fold(La, self._b("a:" + Ls_skip.name), Lu.cout, Ls_skip, self._b("a:" + Lu.name), Lu.cout, Lu, N, h, w)
            self._conv_fwd_fold(Lb, self._b("a:" + La.name), Lb.cin, La, None, 0, None, N, h, w)
            if lvl > 1:           # feeds a transposed convolution: y is materialised; dec1b feeds the head, which folds it
                cur = self._bn_apply(Lb, N, h, w, True)
        return cur

    def _forward_folded(self, x, N, H, W, view=None):
        """Inference forward of the bf16 path (training=False: UNet/model.py:240, inference.py:105): the BatchNorm moving
        statistics are folded into each producer's epilogue (y = act(conv + b) * scale + shift), so every activation
        is written once; max-pool is a plain pool.  Leaves y:dec1b ready for the head."""
        Ls = self.layers

        def fold(L):
            o, c = L.off_stat, L.cout
            return self.FS[o:o + c], self.FB[o:o + c]

        def conv(L, x0, c0, x1, c1, h, w):
            self._cur = L.name
            sc, sh = fold(L)
            y = self._b("y:" + L.name)
            self._call("ub_conv3x3_fwd_affine", x0, c0, x1, c1, self._wptr(L), self.P[L.off_b:L.off_b + L.cout], sc, sh, y, N, h, w, L.cout, 1)
            return y

        L = Ls["enc1a"]
        self._cur = "enc1a"
        sc, sh = fold(L)
        cur = self._b("y:enc1a")
        if view is not None:
            img, img_h, img_w, origins = view
            pitch = img.stride(1)
            plane = img.stride(0) if img.shape[0] > 1 else img.shape[1] * pitch      # the stride of a size-1 dimension is arbitrary
            self._call("ub_conv_first_fwd_affine_tiles", img, origins, img_h, img_w, pitch, plane,
                       self.P[L.off_w:L.off_w + L.n_w], self.P[L.off_b:L.off_b + L.cout], sc, sh, cur, N, H, W, self.number_channels, self.act_code)
        else:
            self._call("ub_conv_first_fwd_affine", x, self.P[L.off_w:L.off_w + L.n_w], self.P[L.off_b:L.off_b + L.cout], sc, sh, cur,
                       N, H, W, self.number_channels, self.act_code)
        for lvl in (1, 2, 3, 4):
            h, w = self._dims(H, W, lvl)
            if lvl > 1:
                La = Ls[f"enc{lvl}a"]
                cur = conv(La, self._b(f"pool{lvl - 1}"), La.cin, None, 0, h, w)
            Lb = Ls[f"enc{lvl}b"]
            cur = conv(Lb, cur, Lb.cin, None, 0, h, w)
            self._call("ub_maxpool2x2_fwd", cur, self._b(f"pool{lvl}"), N, h, w, Lb.cout, self.act_code)
        h, w = self._dims(H, W, 5)
        cur = conv(Ls["bota"], self._b("pool4"), Ls["bota"].cin, None, 0, h, w)
        cur = conv(Ls["botb"], cur, Ls["botb"].cin, None, 0, h, w)
        for lvl in (4, 3, 2, 1):
            hi, wi = self._dims(H, W, lvl + 1)
            h, w = self._dims(H, W, lvl)
            Lu = Ls[f"up{lvl}"]
            self._cur = Lu.name
            sc, sh = fold(Lu)
            u = self._b("y:" + Lu.name)
            self._call("ub_deconv2x2_fwd_affine", cur, Lu.cin, self._wptr(Lu), self.P[Lu.off_b:Lu.off_b + Lu.cout], sc, sh, u, N, hi, wi, Lu.cout)
            La, Lb = Ls[f"dec{lvl}a"], Ls[f"dec{lvl}b"]
            cur = conv(La, self._b(f"y:enc{lvl}b"), Lu.cout, u, Lu.cout, h, w)       # concat [skip, up] (model.py:117)
            cur = conv(Lb, cur, Lb.cin, None, 0, h, w)
        return cur

    def _head_forward(self, N, H, W, training):
        L = self.layers["head"]
        self._cur = "head"
        K = self.number_classes
        P = N * H * W
        a = self._b("a:head")
        if training and self.fold_bn and self.precision == "bf16":
            # dec1b's BatchNorm folded into the 1x1 head (no padding: one bias vector), csrc/fold.cu
            Ld = self.layers["dec1b"]
            L.fold = (Ld, None)
            wf = self._ensure("wfold:head", K * 64, torch.float32)
            bf = self._ensure("bfold:head", _pad8(K), torch.float32)
            sc = self._ensure("fold_s:head", 64, torch.float32)
            sh = self._ensure("fold_t:head", 64, torch.float32)
            mean, rstd = self._bn_vectors(Ld, True)
            gamma, beta = self._affine(Ld)
            self._call("ub_fold_head_weights", self.P[L.off_w:L.off_w + L.n_w], self.P[L.off_b:L.off_b + K], mean, rstd, gamma, beta, wf, bf, sc, sh, K)
            self._call("ub_head_fwd", self._b("a:dec1b"), wf, bf, a, self.partial, P, K, self.act_code)
            self._finalize(L, K, 1, P)
            return a
        self._call("ub_head_fwd", self._b("y:dec1b"), self.P[L.off_w:L.off_w + L.n_w], self.P[L.off_b:L.off_b + K], a,
                   self.partial if training else None, P, K, self.act_code)
        if training:
            self._finalize(L, K, 1, P)
        return a

    def _head_loss(self, N, H, W, training, labels_u8, want_softmax, want_grad):
        L = self.layers["head"]
        K = self.number_classes
        P = N * H * W
        mean, rstd = self._bn_vectors(L, training)
        gamma, beta = self._affine(L)
        sm = self._ensure("softmax", P * K, torch.float32) if want_softmax else None
        inv_denom = 1.0 / (self.global_batch_size * H * W)            # UNet/model.py:213-215
        self._call("ub_head_loss", self._b("a:head"), mean, rstd, gamma, beta, labels_u8, self.class_weights, inv_denom, 1.0 / P, self.label_smoothing,
                   sm, self._b("dlogits") if want_grad else None, self.partial if labels_u8 is not None else None, P, K)
        if labels_u8 is not None:
            self._call("ub_reduce_rows", self.partial, _C.UB_STATS_ROWS, 2, 2, self.metrics, 1.0)
        return sm

    # ------------------------------------------------------------------------------------------------ backward
    def _bn_bwd(self, L, N, h, w, relu):
        """g:<name> holds dL/dy on entry and dL/dz (pre-activation gradient) on exit; fills dgamma/dbeta/dbias in G"""
        self._cur = L.name
        g, a = self._b("g:" + L.name), self._b("a:" + L.name)
        C, M = L.cout, N * h * w
        if getattr(self, "_bwd_inference", False):
            # training=False: BatchNorm is a fixed affine map, dz = gamma * rstd_moving * dy * [a > 0] (zero batch-statistic terms)
            mean, rstd = self._bn_vectors(L, False)
            zero = self._ensure("zeros", 2048, torch.float32)
            zero.zero_()
            self._call("ub_bn_bwd_apply", g, a, mean, rstd, self.P[L.off_gamma:L.off_gamma + C], zero[:C], zero[:C], g, self.partial, M, C,
                       relu, self.act_code)
            return g
        mean, rstd = self._bn_vectors(L, True)
        assert L.off_gamma == L.off_beta + C
        if L.name in self._sums_ready:         # dbeta / dgamma were derived from the consumer's weight gradient (ub_bn_bwd_sums_wgrad)
            self._sums_ready.discard(L.name)
        else:
            if self._red_ready == L.name:      # the dgrad that produced g already accumulated [sum dy | sum dy*xhat]
                src = self.partial_red
            else:
                self._call("ub_bn_bwd_reduce", g, a, mean, rstd, self.partial, M, C, self.act_code)
                src = self.partial
            # [sum dy | sum dy*xhat] lands directly in the flat gradient buffer: beta and gamma segments are adjacent
            self._call("ub_reduce_rows", src, _C.UB_STATS_ROWS, 2 * C, 2 * C, self.G[L.off_beta:L.off_beta + 2 * C], 1.0)
        self._red_ready = None
        dbeta, dgamma = self.G[L.off_beta:L.off_beta + C], self.G[L.off_gamma:L.off_gamma + C]
        self._call("ub_bn_bwd_apply", g, a, mean, rstd, self.P[L.off_gamma:L.off_gamma + C], dbeta, dgamma, g, self.partial, M, C,
                   relu, self.act_code)
        self._call("ub_reduce_rows", self.partial, _C.UB_STATS_ROWS, C, C, self.G[L.off_b:L.off_b + C], 1.0)
        return g

    def _conv_bwd(self, L, x0, x1, N, h, w, dx0, dx1, red=None):
        """red: the BatchNorm'd layer whose dL/dy this conv's dgrad writes last (dx1 of a concat, else dx0): its backward
        sums are accumulated in the dgrad epilogue instead of a separate pass over dy and a (bf16 path)"""
        dz = self._bn_bwd(L, N, h, w, 1)
        dw = self.G[L.off_w:L.off_w + L.n_w]
        c0, c1 = L.c0, L.c1
        if L.fold is not None:          # this layer's forward read pre-BatchNorm activations: so does its weight gradient
            p0, p1 = L.fold
            x0 = self._b("a:" + p0.name) if p0 is not None else x0
            x1 = self._b("a:" + p1.name) if p1 is not None else x1

        # BatchNorm layers whose backward sums follow from THIS layer's weight gradient: the folded sources whose only consumer is this
        # conv (source 0 of a concat is an encoder skip, which also feeds a max-pool: its sums keep their own pass)
        algebra = []
        if self.bn_algebra and L.fold is not None and self.precision == "bf16" and not getattr(self, "_bwd_inference", False):
            p0, p1 = L.fold
            if p0 is not None and c1 == 0:
                algebra.append((p0, 0))
            if p1 is not None:
                algebra.append((p1, c0))
        sums_done = [None]

        def wgrad():
            self._call("ub_conv3x3_wgrad", x0, c0, x1, c1, dz, L.cout, dw, ws, ws.numel(), N, h, w)
            if L.fold is not None:
                # dW = s[ci] * dW_a + t[ci] * (sum of dz over the pixels whose tap neighbour is inside); the total is the bias gradient
                sdz = self._ensure("fold_sdz", 9 * 2048, torch.float32)
                scr = self._ensure("fold_scr", _C.MACROS["UB_BORDER_CHUNKS"] * 8 * 2048, torch.float32)
                self._call("ub_border_sums", dz, self.G[L.off_b:L.off_b + L.cout], sdz, scr, N, h, w, L.cout, self.act_code)
                for p, cb in algebra:
                    pm, pr = self._bn_vectors(p, True)
                    self._call("ub_bn_bwd_sums_wgrad", self.S[L.off_w:L.off_w + L.n_w], _C.UB_BF16, dw, sdz, L.cout, 9, L.cin, cb, p.cout, pm, pr,
                               self.G[p.off_beta:p.off_beta + p.cout], self.G[p.off_gamma:p.off_gamma + p.cout])
                    self._sums_ready.add(p.name)
                if algebra and self.device.type == "cuda":   # the BatchNorm backward of those layers (main stream) waits for the sums, not the fix-up
                    sums_done[0] = torch.cuda.Event()
                    sums_done[0].record(torch.cuda.current_stream(self.device))
                self._call("ub_wgrad_fold_fix", dw, self._b("fold_s:" + L.name), self._b("fold_t:" + L.name), sdz, L.cout, L.cin)

        if self.precision == "bf16":
            ws = self._b("wgrad_ws")
            if not self.overlap_wgrad:
                wgrad()
            # worth it where the K loop is long enough to hide the longer epilogue: layers with >= 128 output channels, and (optional,
            # UB_FUSE_RED64) the 64 -> 64 layers through the row-streaming kernel
            rows_ok = w >= 128 and ((w + 127) // 128 * 128 - w) * 8 <= w          # csrc/conv3_rows.cuh: rows_width_ok
            wide = L.cout >= 128 or (L.cout == 64 and ((self.fuse_bn_reduce_64 >= 2 and rows_ok) or (self.fuse_bn_reduce_64 >= 1 and L.cin == 64)))
            if dx0 is not None and red is not None and self.fuse_bn_reduce and wide and not any(p is red for p, _ in algebra):
                rm, rr = self._bn_vectors(red, True)
                self._call("ub_conv3x3_dgrad_bnred", dz, L.cout, self.WT[L.name], dx0, c0, dx1, c1, N, h, w, self._b("a:" + red.name), rm, rr,
                           self.partial_red)
                self._red_ready = red.name
            elif dx0 is not None:
                self._call("ub_conv3x3_dgrad", dz, L.cout, self.WT[L.name], dx0, c0, dx1, c1, N, h, w)
            if self.overlap_wgrad:
                with torch.cuda.stream(self._fork_side()):
                    wgrad()
                if sums_done[0] is not None:
                    torch.cuda.current_stream(self.device).wait_event(sums_done[0])
        else:
            self._call("ub_check_conv3x3_wgrad", x0, c0, x1, c1, dz, L.cout, dw, N, h, w)
            if dx0 is not None:
                self._call("ub_check_conv3x3", dz, L.cout, None, 0, self.WT[L.name], None, dx0, c0, dx1, c1, N, h, w, 0)

    def _backward(self, x, N, H, W, drop_masks=None, on_layer_done=None, input_grad=None):
        Ls = self.layers
        dm = drop_masks or {}
        K = self.number_classes
        P = N * H * W
        done = on_layer_done or (lambda name: None)
        infer = getattr(self, "_bwd_inference", False)
        if self.overlap_wgrad and self.precision == "bf16":
            self._fork_side()          # the side stream takes part in this step (and in its graph capture) from here on
        # ---- head: BN backward + relu mask + 1x1 dgrad/wgrad
        L = Ls["head"]
        self._cur = "head"
        mean, rstd = self._bn_vectors(L, not infer)
        dl, a = self._b("dlogits"), self._b("a:head")
        head_x = self._b("a:dec1b") if L.fold is not None else self._b("y:dec1b")       # folded head: its weight gradient reads `a` too
        dbeta, dgamma = self.G[L.off_beta:L.off_beta + K], self.G[L.off_gamma:L.off_gamma + K]
        if infer:
            dbeta.zero_()
            dgamma.zero_()
        else:
            self._call("ub_head_bwd_reduce", dl, a, mean, rstd, self.partial, P, K)
            self._call("ub_reduce_rows", self.partial, _C.UB_STATS_ROWS, 2 * K, K, dbeta, 1.0)
            self._call("ub_reduce_rows", self.partial[K:], _C.UB_STATS_ROWS, 2 * K, K, dgamma, 1.0)
        if self.fuse_bn_reduce_ew and self.precision == "bf16" and not infer and K <= 4:
            # the head's dgrad writes dL/dy of dec1b: that layer's BatchNorm-backward sums come out of the same pass
            Ld = Ls["dec1b"]
            rm, rr = self._bn_vectors(Ld, True)
            self._call("ub_head_bwd_apply_bnred", dl, a, head_x, self.P[L.off_w:L.off_w + L.n_w], mean, rstd,
                       self.P[L.off_gamma:L.off_gamma + K], dbeta, dgamma, self._b("g:dec1b"), self.partial, P, K, self.act_code,
                       self._b("a:dec1b"), rm, rr, self.partial_red)
            self._red_ready = "dec1b"
        else:
            self._call("ub_head_bwd_apply", dl, a, head_x, self.P[L.off_w:L.off_w + L.n_w], mean, rstd,
                       self.P[L.off_gamma:L.off_gamma + K], dbeta, dgamma, self._b("g:dec1b"), self.partial, P, K, self.act_code)
        ncomp = K * 64 + K
        self._call("ub_reduce_rows", self.partial, _C.UB_STATS_ROWS, ncomp, K * 64, self.G[L.off_w:L.off_w + K * 64], 1.0)
        self._call("ub_reduce_rows", self.partial[K * 64:], _C.UB_STATS_ROWS, ncomp, K, self.G[L.off_b:L.off_b + K], 1.0)
        if L.fold is not None and self.bn_algebra and not infer and not (self.fuse_bn_reduce_ew and K <= 4):
            # dec1b's BatchNorm-backward sums from the head's weight gradient on `a` and its bias gradient (a 1x1 conv: one tap, no border)
            Ld = Ls["dec1b"]
            pm, pr = self._bn_vectors(Ld, True)
            self._call("ub_bn_bwd_sums_wgrad", self.P[L.off_w:L.off_w + K * 64], _C.UB_F32, self.G[L.off_w:L.off_w + K * 64],
                       self.G[L.off_b:L.off_b + K], K, 1, 64, 0, 64, pm, pr, self.G[Ld.off_beta:Ld.off_beta + 64], self.G[Ld.off_gamma:Ld.off_gamma + 64])
            self._sums_ready.add("dec1b")
        if L.fold is not None:          # dW[k][c] = s[c] dW_a[k][c] + t[c] db[k]
            self._call("ub_head_wgrad_fold_fix", self.G[L.off_w:L.off_w + K * 64], self.G[L.off_b:L.off_b + K], self._b("fold_s:head"),
                       self._b("fold_t:head"), K)
        done("head")
        # ---- decoder
        for lvl in (1, 2, 3, 4):
            h, w = self._dims(H, W, lvl)
            hi, wi = self._dims(H, W, lvl + 1)
            La, Lb, Lu = Ls[f"dec{lvl}a"], Ls[f"dec{lvl}b"], Ls[f"up{lvl}"]
            self._conv_bwd(Lb, self._b("y:" + La.name), None, N, h, w, self._b("g:" + La.name), None, red=La)
            done(Lb.name)
            self._conv_bwd(La, self._b(f"y:enc{lvl}b"), self._b("y:" + Lu.name), N, h, w, self._b(f"gskip{lvl}"), self._b("g:" + Lu.name), red=Lu)
            done(La.name)
            prev = Ls[f"dec{lvl + 1}b"] if lvl < 4 else Ls["botb"]
            dz = self._bn_bwd(Lu, N, h, w, 0)
            dw = self.G[Lu.off_w:Lu.off_w + Lu.n_w]
            xin = self._b("y:" + prev.name)
            if self.precision == "bf16":
                ws = self._b("wgrad_ws")
                if not self.overlap_wgrad:
                    self._call("ub_deconv2x2_wgrad", xin, Lu.cin, dz, Lu.cout, dw, ws, ws.numel(), N, hi, wi)
                if self.fuse_bn_reduce and self.fuse_bn_reduce_deconv and lvl < 4 and not infer:
                    rm, rr = self._bn_vectors(prev, True)
                    self._call("ub_deconv2x2_dgrad_bnred", dz, Lu.cout, self.WT[Lu.name], self._b("g:" + prev.name), Lu.cin, N, hi, wi,
                               self._b("a:" + prev.name), rm, rr, self.partial_red)
                    self._red_ready = prev.name
                else:
                    self._call("ub_deconv2x2_dgrad", dz, Lu.cout, self.WT[Lu.name], self._b("g:" + prev.name), Lu.cin, N, hi, wi)
                if self.overlap_wgrad:
                    with torch.cuda.stream(self._fork_side()):
                        self._call("ub_deconv2x2_wgrad", xin, Lu.cin, dz, Lu.cout, dw, ws, ws.numel(), N, hi, wi)
            else:
                self._call("ub_check_deconv2x2_wgrad", xin, dz, dw, N, hi, wi, Lu.cin, Lu.cout)
                self._call("ub_check_deconv2x2_dgrad", dz, self._wptr(Lu), self._b("g:" + prev.name), N, hi, wi, Lu.cin, Lu.cout)
            done(Lu.name)
        # ---- bottleneck
        h, w = self._dims(H, W, 5)
        La, Lb = Ls["bota"], Ls["botb"]
        if "dropb" in dm:
            g = self._b("g:botb")
            self._call("ub_dropout_bwd", g, dm["dropb"], g, N * h * w * Lb.cout, self.act_code)
        self._conv_bwd(Lb, self._b("y:bota"), None, N, h, w, self._b("g:bota"), None, red=La)
        done("botb")
        self._conv_bwd(La, self._b("pool4"), None, N, h, w, self._b("gpool4"), None)
        done("bota")
        # ---- encoder
        for lvl in (4, 3, 2, 1):
            h, w = self._dims(H, W, lvl)
            La, Lb = Ls[f"enc{lvl}a"], Ls[f"enc{lvl}b"]
            if self.fuse_bn_reduce_ew and self.precision == "bf16" and not infer:
                # the pass that assembles dL/dy of enc<l>b also accumulates that layer's BatchNorm-backward sums
                rm, rr = self._bn_vectors(Lb, True)
                self._cur = Lb.name
                self._call("ub_maxpool2x2_bwd_add_bnred", self._b(f"gpool{lvl}"), self._b(f"idx{lvl}"), self._b(f"gskip{lvl}"),
                           dm.get("drop4") if lvl == 4 else None, self._b("g:" + Lb.name), N, h, w, Lb.cout, self._b("a:" + Lb.name), rm, rr,
                           self.partial_red, self.act_code)
                self._red_ready = Lb.name
            else:
                self._call("ub_maxpool2x2_bwd_add", self._b(f"gpool{lvl}"), self._b(f"idx{lvl}"), self._b(f"gskip{lvl}"),
                           dm.get("drop4") if lvl == 4 else None, self._b("g:" + Lb.name), N, h, w, Lb.cout, self.act_code)
            self._conv_bwd(Lb, self._b("y:" + La.name), None, N, h, w, self._b("g:" + La.name), None, red=La)
            done(Lb.name)
            if lvl > 1:
                self._conv_bwd(La, self._b(f"pool{lvl - 1}"), None, N, h, w, self._b(f"gpool{lvl - 1}"), None)
            else:
                dz = self._bn_bwd(La, N, h, w, 1)
                self._call("ub_conv_first_wgrad", x, dz, self.G[La.off_w:La.off_w + La.n_w], self._b("first_ws"), N, H, W,
                           self.number_channels, self.act_code)
                if input_grad is not None:
                    self._call("ub_conv_first_dgrad", dz, self.P[La.off_w:La.off_w + La.n_w], input_grad, N, H, W, self.number_channels,
                               self.act_code)
            done(La.name)
        self._join_side()

    # ------------------------------------------------------------------------------------------------ optimizer
    def _lr_t(self):
        t = self.step_count
        return self.learning_rate * math.sqrt(1.0 - ADAM_B2 ** t) / (1.0 - ADAM_B1 ** t)

    def _upload_lr_t(self):
        """this step's bias-corrected learning rate -> device scalar (pinned ring: the host may run a few steps ahead)"""
        if self._lr_ring is None:
            self._lr_ring = torch.zeros(256, dtype=torch.float32).pin_memory()
            self._lr_dev = torch.zeros(1, dtype=torch.float32, device=self.device)
        i = self.step_count % 256
        self._lr_ring[i] = self._lr_t()
        self._lr_dev.copy_(self._lr_ring[i:i + 1], non_blocking=True)

    def _adam(self, lo=0, hi=None, lr_on_device=False):
        self._cur = "optimizer"
        hi = self.n_flat if hi is None else hi
        if lr_on_device:
            self._call("ub_adam_dev", self.P[lo:hi], self.G[lo:hi], self.M[lo:hi], self.V[lo:hi],
                       self.S[lo:hi] if self.S is not None else None, hi - lo, self._lr_dev, ADAM_B1, ADAM_B2, ADAM_EPS, 1.0)
            return
        self._call("ub_adam", self.P[lo:hi], self.G[lo:hi], self.M[lo:hi], self.V[lo:hi],
                   self.S[lo:hi] if self.S is not None else None, hi - lo, self._lr_t(), ADAM_B1, ADAM_B2, ADAM_EPS, 1.0)

    # ------------------------------------------------------------------------------------------------ inputs
    def _prep_images(self, images):
        x = images if torch.is_tensor(images) else torch.as_tensor(np.asarray(images))
        if x.dim() != 4 or x.shape[1] != self.number_channels:
            raise IOError(f"images must be [N,{self.number_channels},H,W], got {tuple(x.shape)}")
        x = x.to(device=self.device, dtype=torch.float32, non_blocking=True).contiguous()
        return x

    def _prep_labels(self, labels, N, H, W):
        """one-hot int32 [N,H,W,K] (reference contract, imagereader.py:353-355) or uint8/int class index [N,H,W]"""
        t = labels if torch.is_tensor(labels) else torch.as_tensor(np.asarray(labels))
        t = t.to(self.device, non_blocking=True)
        K = self.number_classes
        if t.dim() == 4:
            if tuple(t.shape) != (N, H, W, K):
                raise IOError(f"labels must be [N,H,W,{K}] one-hot, got {tuple(t.shape)}")
            idx = self._ensure("labels_u8", N * H * W, torch.uint8)
            self._call("ub_onehot_to_index", t.to(torch.int32).contiguous(), idx, N * H * W, K)
            return idx
        if tuple(t.shape) != (N, H, W):
            raise IOError(f"labels must be [N,H,W] class indices, got {tuple(t.shape)}")
        return t.to(torch.uint8).contiguous()

    def normalize_batch(self, raw, src_code=None, slot=""):
        """per-tile, per-channel z-score on the device (UNet/imagereader.py:300, :33-66): raw [N,C,H,W] uint8 / uint16
        (as int16 bits) / float32 device tensor -> float32 NCHW in a persistent buffer (one per `slot`)"""
        N, C, H, W = raw.shape
        if src_code is None:
            src_code = {torch.uint8: 0, torch.int16: 1, torch.float32: 2}.get(raw.dtype, 1 if str(raw.dtype) == "torch.uint16" else None)
            if src_code is None:
                raise IOError(f"unsupported pixel dtype {raw.dtype}")
        out = self._ensure("x_norm" + str(slot), N * C * H * W, torch.float32)[:N * C * H * W].view(N, C, H, W)
        scratch = self._ensure("zscore_scratch" + str(slot), N * C * _C.UB_ZSCORE_BLOCKS * 2, torch.float64)
        self._cur = "input"
        self._call("ub_zscore", raw.contiguous(), src_code, out, scratch, N * C, H * W)
        return out

    def _make_drop_masks(self, N, H, W):
        b = self._BASELINE_FEATURE_DEPTH
        n4 = N * (H // 8) * (W // 8) * 8 * b
        nb = N * (H // 16) * (W // 16) * 16 * b
        m4 = self._ensure("drop4", _pad8(n4) + 16, torch.uint8)
        mb = self._ensure("dropb", _pad8(nb) + 16, torch.uint8)
        seed = getattr(self, "dropout_seed", 0x5EED) + (self.dist.rank if self.dist is not None else 0) * 7919
        off = self.step_count * (1 << 24)
        self._call("ub_dropout_mask", m4, (n4 + 15) // 16 * 16, seed, off)
        self._call("ub_dropout_mask", mb, (nb + 15) // 16 * 16, seed + 1, off)
        return {"drop4": m4, "dropb": mb}

    def _import_drop_masks(self, masks):
        """oracle-style masks (NCHW, {0,1}) -> NHWC uint8 device tensors"""
        out = {}
        for k, v in masks.items():
            t = torch.as_tensor(np.asarray(v)).to(torch.uint8)
            out[k] = t.permute(0, 2, 3, 1).contiguous().to(self.device)
        return out

    # ------------------------------------------------------------------------------------------------ steps
    def train_step(self, inputs, labels=None, *, dropout_masks=None, apply_update=True, keep_softmax=False):
        """UNet/model.py:204-228.  Accepts the reference tuple (images, labels, loss_metric, accuracy_metric) or
        (images, labels).  Returns the loss as a 0-d device tensor (no host sync).
        dropout_masks: None -> fresh Philox masks; {} / False -> no dropout; dict of NCHW {0,1} arrays -> injected."""
        loss_metric = acc_metric = None
        if labels is None:
            if len(inputs) == 4:
                images, labels, loss_metric, acc_metric = inputs
            else:
                images, labels = inputs
        else:
            images = inputs
        x = self._prep_images(images)
        N, _, H, W = x.shape
        lab = self._prep_labels(labels, N, H, W)
        self.step_count += 1
        if dropout_masks is None:
            dm = self._make_drop_masks(N, H, W)
        elif not dropout_masks:
            dm = {}
        else:
            dm = self._import_drop_masks(dropout_masks)
        graphable = (self.use_graph and dropout_masks is None and apply_update and not keep_softmax and self.profile is None
                     and self.precision == "bf16")
        if graphable:
            self._train_step_graph(x, lab, N, H, W, dm)
        else:
            self._step_body(x, lab, N, H, W, dm, keep_softmax, apply_update, False)
        if apply_update:
            self._inference_stale = True
            self._radius_cache = None
        loss = self.metrics[0]
        if loss_metric is not None:
            loss_metric.update_state(loss)
        if acc_metric is not None:
            acc_metric.update_state(self.metrics[1])
        return loss

    def _step_body(self, x, lab, N, H, W, dm, keep_softmax, apply_update, lr_on_device):
        """forward, loss, backward (+ bucketed all-reduce), Adam, dgrad repack: every launch of one optimisation step"""
        self._forward(x, N, H, W, True, dm)
        self._head_forward(N, H, W, True)
        self._head_loss(N, H, W, True, lab, keep_softmax, True)
        if self.dist is not None and self.dist.world_size > 1:
            self.dist.begin_step(self)
            self._backward(x, N, H, W, dm, on_layer_done=lambda name: self.dist.layer_done(self, name))
            self.dist.finish_step(self)
        else:
            self._backward(x, N, H, W, dm)
        if apply_update:
            self._adam(lr_on_device=lr_on_device)
            self._repack_dgrad()

    def _train_step_graph(self, x, lab, N, H, W, dm):
        """The step is ~230 kernel launches and ~75 memsets; replaying it from a CUDA graph removes the launch gaps (measured
        1.07 ms of 24.8 ms, tools/graph_probe.py).  What varies per step stays outside the graph: the dropout masks (already
        generated into persistent buffers), the bias-corrected learning rate (device scalar) and the inputs (copied into
        persistent staging tensors).  First call per shape runs eagerly (allocations, attribute calls), the second captures."""
        self._upload_lr_t()
        key = (N, H, W)
        st = self._graphs.get(key)
        if st is None:
            self._graphs[key] = {"graph": None}
            self._step_body(x, lab, N, H, W, dm, False, True, True)
            return
        xin = self._ensure("graph_x", x.numel(), torch.float32)[:x.numel()].view_as(x)
        lin = self._ensure("graph_lab", lab.numel(), torch.uint8)[:lab.numel()].view_as(lab)
        xin.copy_(x)
        lin.copy_(lab)
        if st["graph"] is None:
            g = torch.cuda.CUDAGraph()
            l0 = self.launches
            with torch.cuda.graph(g):
                self._step_body(xin, lin, N, H, W, dm, False, True, True)
            st["graph"] = g
            st["launches"] = self.launches - l0
        else:
            self.launches += st["launches"]
        st["graph"].replay()

    def test_step(self, inputs, labels=None):
        """UNet/model.py:237-250: training=False (moving statistics, no dropout)."""
        loss_metric = acc_metric = None
        if labels is None:
            if len(inputs) == 4:
                images, labels, loss_metric, acc_metric = inputs
            else:
                images, labels = inputs
        else:
            images = inputs
        x = self._prep_images(images)
        N, _, H, W = x.shape
        lab = self._prep_labels(labels, N, H, W)
        self._forward(x, N, H, W, False)
        self._head_forward(N, H, W, False)
        self._head_loss(N, H, W, False, lab, False, False)
        loss = self.metrics[0].clone()
        if loss_metric is not None:
            loss_metric.update_state(loss)
        if acc_metric is not None:
            acc_metric.update_state(self.metrics[1].clone())
        return loss

    def dist_train_step(self, dist_strategy, inputs):
        """UNet/model.py:230-235: per-replica step + SUM of the per-replica losses (each already divided by the
        global batch).  One process per GPU here: `dist_strategy` reduces the scalar across ranks."""
        loss = self.train_step(inputs)
        return dist_strategy.reduce_sum(loss) if dist_strategy is not None else loss

    def dist_test_step(self, dist_strategy, inputs):
        loss = self.test_step(inputs)
        return dist_strategy.reduce_sum(loss) if dist_strategy is not None else loss

    def forward_softmax(self, images, training=False, dropout_masks=None):
        """model(images, training=...) of the reference: NCHW float32 -> softmax NHWC [N,H,W,K] (device tensor)."""
        x = self._prep_images(images)
        N, _, H, W = x.shape
        dm = self._import_drop_masks(dropout_masks) if (training and dropout_masks) else {}
        self._forward(x, N, H, W, training, dm)
        self._head_forward(N, H, W, training)
        sm = self._head_loss(N, H, W, training, None, True, False)
        return sm[:N * H * W * self.number_classes].view(N, H, W, self.number_classes)

    def _model_call(self, batch_data):
        """get_keras_model()(batch) contract of UNet/inference.py:105: numpy in, array-like softmax out."""
        return self.forward_softmax(batch_data, training=False).cpu().numpy()

    def predict_tiles_into(self, x, geo, mask, mask_ld):
        """Inference fast path: forward a batch of equal-sized NCHW fp32 device tiles and write the argmax of each
        tile's zone of responsibility straight into the uint8 device mask (UNet/inference.py:105-129 without the
        softmax round trip).  geo: int32 device tensor [N,6] = (cy0, cy1, cx0, cx1, dst_y, dst_x) per tile."""
        N, _, H, W = x.shape
        self._forward(x, N, H, W, False)
        L = self.layers["head"]
        K = self.number_classes
        o = L.off_stat
        self._cur = "head"
        self._call("ub_head_argmax", self._b("y:dec1b"), self.P[L.off_w:L.off_w + L.n_w], self.P[L.off_b:L.off_b + K],
                   self.FS[o:o + K], self.FB[o:o + K], K, N, H, W, geo, mask, mask_ld, None, self.act_code)

    def predict_tiles_from(self, img, img_h, img_w, origins, n, h, w, geo, mask, mask_ld):
        """The same for `n` tiles of h x w read IN PLACE from the resident normalised image `img` [C, rows, pitch] (fp32): tile i
        starts at (origins[i, 0], origins[i, 1]); rows / columns at or beyond the unpadded extent (img_h, img_w) are mirrored
        (the reflect padding to a multiple of 16, UNet/inference.py:46).  No per-tile copy, no padded copy of the image."""
        if self.precision != "bf16":
            # fp32 check mode: materialise the tiles (mirror indices computed here), then the plain path
            ys = torch.arange(h, device=img.device)
            xs = torch.arange(w, device=img.device)
            tiles = []
            for oy, ox in origins[:n].tolist():
                yy, xx = ys + oy, xs + ox
                yy = torch.where(yy < img_h, yy, 2 * (img_h - 1) - yy)
                xx = torch.where(xx < img_w, xx, 2 * (img_w - 1) - xx)
                tiles.append(img[:, yy][:, :, xx])
            return self.predict_tiles_into(torch.stack(tiles).contiguous(), geo, mask, mask_ld)
        self._forward(None, n, h, w, False, view=(img, int(img_h), int(img_w), origins))
        L = self.layers["head"]
        K = self.number_classes
        o = L.off_stat
        self._cur = "head"
        self._call("ub_head_argmax", self._b("y:dec1b"), self.P[L.off_w:L.off_w + L.n_w], self.P[L.off_b:L.off_b + K],
                   self.FS[o:o + K], self.FB[o:o + K], K, n, h, w, geo, mask, mask_ld, None, self.act_code)

    # ------------------------------------------------------------------------------------------------ checkpoint
    def state_dict(self):
        return {"P": self.P.cpu(), "M": self.M.cpu(), "V": self.V.cpu(), "MM": self.MM.cpu(), "MV": self.MV.cpu(),
                "step": self.step_count, "number_classes": self.number_classes, "number_channels": self.number_channels,
                "learning_rate": self.learning_rate}

    def _blocks(self):
        return list(self.layers), [L.kind for L in self.layers.values()]

    def save_checkpoint(self, checkpoint_filepath, format="tf"):
        """Counterpart of tf.train.Checkpoint(optimizer, model).write(prefix) (UNet/train.py:96, :181-184).
        format="tf": the same two files the reference leaves behind, <prefix>.index + <prefix>.data-00000-of-00001 (a
        TensorBundle with the TF2 object-graph keys, unetb200/tfcheckpoint.py); format="native": one torch.save file."""
        if format == "native":
            torch.save(self.state_dict(), checkpoint_filepath)
            return
        if format != "tf":
            raise ValueError("format must be 'tf' or 'native'")
        from . import tfcheckpoint
        names, kinds = self._blocks()
        tfcheckpoint.save_unet(checkpoint_filepath, names, kinds, self.export_params(), self.export_flat(self.M), self.export_flat(self.V),
                               step=self.step_count, learning_rate=self.learning_rate, beta_1=ADAM_B1, beta_2=ADAM_B2)

    def load_checkpoint(self, checkpoint_filepath: str):     # UNet/model.py:81-83 (expect_partial: optimizer slots optional)
        import os
        if os.path.exists(checkpoint_filepath + ".index"):
            from . import tfcheckpoint
            names, kinds = self._blocks()
            ck = tfcheckpoint.load_unet(checkpoint_filepath, names, kinds)
            k = ck["params"]["head/kernel"].shape[-1]
            c = ck["params"]["enc1a/kernel"].shape[2]
            if k != self.number_classes or c != self.number_channels:
                raise IOError("checkpoint was written for a different number_classes / number_channels")
            if ck["adam_m"] is not None:
                self.M.copy_(self._import_flat(ck["adam_m"], torch.zeros(self.n_flat)))
                self.V.copy_(self._import_flat(ck["adam_v"], torch.zeros(self.n_flat)))
                self.step_count = int(ck["step"] or 0)
            self.load_oracle_params(ck["params"])
            return
        sd = torch.load(checkpoint_filepath, map_location="cpu")
        if sd["number_classes"] != self.number_classes or sd["number_channels"] != self.number_channels:
            raise IOError("checkpoint was written for a different number_classes / number_channels")
        self.P.copy_(sd["P"])
        self.MM.copy_(sd["MM"])
        self.MV.copy_(sd["MV"])
        if "M" in sd:
            self.M.copy_(sd["M"])
            self.V.copy_(sd["V"])
            self.step_count = int(sd.get("step", 0))
        self._weights_changed()
