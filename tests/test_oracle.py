"""CPU (-m "not gpu"): pins the oracle.  The reference has no tests / golden vectors and TensorFlow is not installable
here (parity unpinned, see oracle/unet_oracle.py), so the oracle is pinned by
  (1) two independent derivations agreeing in fp64 (torch autograd vs hand-written numpy forward+backward),
  (2) finite differences,
  (3) invariants of the reference graph,
  (4) the committed golden fixtures reproducing (tests/golden/make_golden.py),
  (5) the numpy restatements of the reference's host code (z-score, one-hot, tiling) on hand-checked cases.
"""
import os

import numpy as np
import pytest
import torch

from oracle import unet_numpy as ON
from oracle import unet_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _inputs(N, C, H, W, K, seed):
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(N, C, H, W))
    lab = rng.integers(0, K, size=(N, H, W))
    oh = np.eye(K, dtype=np.int32)[lab]
    dm = {"drop4": rng.integers(0, 2, size=(N, 8 * 8, H // 8, W // 8)), "dropb": rng.integers(0, 2, size=(N, 16 * 8, H // 16, W // 16))}
    return x, oh, dm


@pytest.fixture(scope="module")
def small():
    """base width 8 (same topology, 8x fewer channels) keeps the fp64 numpy derivation fast"""
    N, C, H, W, K = 2, 3, 32, 48, 4
    p = O.init_params(C, K, seed=1, base=8, randomize_affine=True)
    x, oh, dm = _inputs(N, C, H, W, K, 1)
    r = O.train_step_grads(p, torch.tensor(x), torch.tensor(oh), N, {k: torch.tensor(v) for k, v in dm.items()})
    return dict(p=p, x=x, oh=oh, dm=dm, r=r, N=N, K=K)


def test_two_derivations_agree(small):
    m = ON.ManualUNet({k: v.numpy() for k, v in small["p"].items()})
    sm = m.forward(small["x"], small["dm"])
    loss = m.loss(small["oh"], small["N"])
    g = m.backward()
    r = small["r"]
    assert abs(loss - float(r["loss"])) < 1e-12
    assert np.abs(sm - r["softmax"].numpy()).max() < 1e-12
    assert set(g) == set(r["grads"])
    for k, v in r["grads"].items():
        ref = v.numpy()
        scale = max(np.abs(ref).max(), 1e-9)
        if k.startswith("up") and k.endswith("/bias"):       # analytically zero
            assert np.abs(g[k]).max() < 1e-12 and np.abs(ref).max() < 1e-12
        else:
            assert np.abs(g[k] - ref).max() / scale < 1e-9, k


def test_finite_differences(small):
    p, x, oh, dm = small["p"], torch.tensor(small["x"]), torch.tensor(small["oh"]), {k: torch.tensor(v) for k, v in small["dm"].items()}
    rng = np.random.default_rng(0)

    def loss_of(params):
        _, logits = O.forward(params, x, True, dm)
        return float(O.loss_and_accuracy(logits, oh, small["N"])[0])

    for name in ("enc1a/kernel", "enc3b/gamma", "botb/kernel", "up2/kernel", "dec1a/bias", "head/kernel", "head/beta"):
        g = small["r"]["grads"][name]
        idx = tuple(int(rng.integers(0, s)) for s in g.shape)
        eps = 1e-6
        pp = {k: v.clone() for k, v in p.items()}
        pp[name][idx] += eps
        up = loss_of(pp)
        pp[name][idx] -= 2 * eps
        dn = loss_of(pp)
        fd = (up - dn) / (2 * eps)
        assert abs(fd - float(g[idx])) < 1e-6 * max(1.0, abs(fd)) + 5e-8, (name, fd, float(g[idx]))


def test_invariants(small):
    r, K = small["r"], small["K"]
    sm = r["softmax"].numpy()
    assert np.allclose(sm.sum(-1), 1.0, atol=1e-12)
    assert 0.5 * np.log(K) < float(r["loss"]) < 3 * np.log(K)
    taps = {}
    O.forward(small["p"], torch.tensor(small["x"]), True, {k: torch.tensor(v) for k, v in small["dm"].items()}, None, taps)
    p0 = O.init_params(3, K, seed=1, base=8)                 # gamma 1 / beta 0: BN output is standardised
    taps0 = {}
    O.forward(p0, torch.tensor(small["x"]), True, None, None, taps0)
    for n, t in taps0.items():
        if n.endswith("/out"):
            assert abs(float(t.mean(dim=(0, 2, 3)).abs().max())) < 1e-9
            v = t.var(dim=(0, 2, 3), unbiased=False)
            assert float(v.max()) <= 1.0 + 1e-9             # var/(var+eps) <= 1


def test_conditioning_is_identity_on_own_pattern(small):
    """injecting the oracle's OWN activation pattern and pool routing must not change anything"""
    p, x, oh = small["p"], torch.tensor(small["x"]), torch.tensor(small["oh"])
    dm = {k: torch.tensor(v) for k, v in small["dm"].items()}
    taps = {}
    O.forward(p, x, True, dm, None, taps)
    relu = {n[:-4]: taps[n] > 0 for n in taps if n.endswith("/act") and not n.startswith("up")}
    pool = {}
    for lvl in (1, 2, 3, 4):
        y = taps[f"enc{lvl}b/out"]
        if lvl == 4:
            y = y * dm["drop4"] * 2.0
        n, c, h, w = y.shape
        win = y.reshape(n, c, h // 2, 2, w // 2, 2).permute(0, 1, 2, 4, 3, 5).reshape(n, c, h // 2, w // 2, 4)
        pool[f"pool{lvl}"] = win.argmax(-1)
    rc = O.train_step_grads(p, x, oh, small["N"], dm, relu_masks=relu, pool_idx=pool)
    assert abs(float(rc["loss"]) - float(small["r"]["loss"])) < 1e-13
    for k, v in small["r"]["grads"].items():
        assert float((rc["grads"][k] - v).abs().max()) < 1e-12, k


def test_keras_adam_formula():
    """App. A.6: eps is added to sqrt(v) without bias-correcting v"""
    p = {"w/kernel": torch.tensor([1.0, -2.0], dtype=torch.float64)}
    opt = O.KerasAdam(p, 0.1)
    g = {"w/kernel": torch.tensor([0.5, -0.25], dtype=torch.float64)}
    opt.apply(p, g)
    lr_t = 0.1 * np.sqrt(1 - 0.999) / (1 - 0.9)
    m, v = 0.1 * np.array([0.5, -0.25]), 0.001 * np.array([0.25, 0.0625])
    assert np.allclose(p["w/kernel"].numpy(), np.array([1.0, -2.0]) - lr_t * m / (np.sqrt(v) + 1e-7), rtol=0, atol=1e-15)


def test_bn_moving_statistics_use_unbiased_variance():
    p = O.init_params(1, 2, seed=0, base=8)
    x = torch.tensor(np.random.default_rng(0).normal(size=(2, 1, 16, 16)))
    ns = {}
    taps = {}
    O.forward(p, x, True, None, ns, taps)
    a = taps["enc1a/act"]
    n = a.numel() // a.shape[1]
    assert torch.allclose(ns["enc1a/moving_var"], 0.99 * torch.ones(8, dtype=torch.float64) + 0.01 * a.var(dim=(0, 2, 3), unbiased=False) * n / (n - 1))
    assert torch.allclose(ns["enc1a/moving_mean"], 0.01 * a.mean(dim=(0, 2, 3)))


@pytest.mark.parametrize("name", ["graph_c1_k2", "graph_c3_k8"])
def test_golden_fixture_reproduces(name):
    g = dict(np.load(os.path.join(GOLD, name + ".npz")))
    N, C, H, W, K = (int(g[k]) for k in "NCHWK")
    p = O.init_params(C, K, seed=int(g["seed"]), base=64, randomize_affine=True)
    d4 = np.unpackbits(g["drop4"])[:N * 512 * (H // 8) * (W // 8)].reshape(N, 512, H // 8, W // 8)
    db = np.unpackbits(g["dropb"])[:N * 1024 * (H // 16) * (W // 16)].reshape(N, 1024, H // 16, W // 16)
    _, logits = O.forward(p, torch.tensor(g["x"], dtype=torch.float64), True, {"drop4": torch.tensor(d4), "dropb": torch.tensor(db)})
    loss, acc = O.loss_and_accuracy(logits, torch.tensor(np.eye(K, dtype=np.int32)[g["labels"]]), int(g["gb"]))
    assert abs(float(loss) - float(g["loss"])) < 1e-12
    assert abs(float(acc) - float(g["acc"])) < 1e-12
    assert np.abs(torch.softmax(logits, -1).numpy() - g["softmax"]).max() < 1e-6      # fixture stores float32


# ------------------------------------------------------------------------------------------------ host restatements
def test_zscore_matches_reference_formula():
    h = dict(np.load(os.path.join(GOLD, "host.npz")))
    z = O.zscore_normalize(h["img"])
    assert np.array_equal(z, h["zscore"])
    img = h["img"].astype(np.float32)
    for c in range(3):
        std, mu = np.std(img[c]), np.mean(img[c])
        want = (img[c] - mu) if std <= 1.0 else (img[c] - mu) / std       # UNet/imagereader.py:44-49
        assert np.array_equal(z[c], want)
    assert np.all(z[2] == 0)                                                  # constant plane: std <= 1 -> x - mean
    hwc = O.zscore_normalize(np.transpose(h["img"], (1, 2, 0)), channels_first=False)
    assert np.array_equal(np.transpose(hwc, (2, 0, 1)), z)
    with pytest.raises(IOError):
        O.zscore_normalize(np.zeros((2, 2, 2, 2)))


def test_one_hot_and_label_range():
    h = dict(np.load(os.path.join(GOLD, "host.npz")))
    oh = O.one_hot(h["lab"], 4)
    assert np.array_equal(oh, h["onehot"]) and oh.dtype == np.int32
    assert np.array_equal(oh.argmax(-1), h["lab"]) and np.all(oh.sum(-1) == 1)
    with pytest.raises(IndexError):                                          # UNet/imagereader.py:307-311
        O.one_hot(h["lab"], 3)


def test_narrow_mask_dtype_quirk():
    assert O.narrow_mask_dtype(np.array([[0, 255]], dtype=np.int32)).dtype == np.uint8
    assert O.narrow_mask_dtype(np.array([[0, 256]], dtype=np.int32)).dtype == np.uint16
    assert O.narrow_mask_dtype(np.array([[0, 65535]], dtype=np.int32)).dtype == np.uint16
    assert O.narrow_mask_dtype(np.array([[0, 65536]], dtype=np.int32)).dtype == np.int32     # Q13: matches no branch
    assert O.narrow_mask_dtype(np.array([[0, 70000]], dtype=np.int32)).dtype == np.int32


def test_tile_plan_geometry():
    h = dict(np.load(os.path.join(GOLD, "host.npz")))
    plan = O.tile_plan(2000, 2512, 1024, 96)
    keys = sorted(plan[0])
    assert [str(k) for k in h["plan_keys"]] == keys
    assert np.array_equal(np.array([[t[k] for k in keys] for t in plan]), h["plan"])
    zone = 1024 - 192
    assert len(plan) == -(-2000 // zone) * -(-2512 // zone)
    cover = np.zeros((2000, 2512), dtype=np.int32)
    for t in plan:
        assert (t["y_end"] - t["y_st"]) % 16 == 0 and (t["x_end"] - t["x_st"]) % 16 == 0
        assert t["y_end"] - t["y_st"] <= 1024 and t["x_end"] - t["x_st"] <= 1024
        cover[t["y_st_z"]:t["y_end_z"], t["x_st_z"]:t["x_end_z"]] += 1
    assert cover.min() >= 1
    # config 5: 20000^2, radius 96 -> 25 x 25 = 625 tiles, 484 of them full 1024 x 1024
    big = O.tile_plan(20000, 20000, 1024, 96)
    assert len(big) == 625
    assert sum(1 for t in big if t["y_end"] - t["y_st"] == 1024 and t["x_end"] - t["x_st"] == 1024) == 484


def test_tiled_inference_equals_whole_image_with_context_free_model():
    """with a per-pixel model (no spatial context) tiling must reproduce the whole-image mask exactly, including the
    reflect padding of sizes that are not multiples of 16 and the last-writer-wins rows/columns (Q10, Q11)"""
    rng = np.random.default_rng(0)
    img = rng.normal(size=(150, 219, 1)).astype(np.float32)

    def model_fn(batch):
        assert batch.shape[2] % 16 == 0 and batch.shape[3] % 16 == 0
        v = batch[0, 0]
        return np.stack([np.sin(3 * v), np.cos(2 * v), v * 0.1], axis=-1)[None]

    whole = O.inference_whole(img, model_fn)
    tiled = O.inference_tiling(img, model_fn, tile_size=96, radius=16)
    assert whole.shape == (150, 219) and whole.dtype == np.int32
    assert np.array_equal(whole, tiled)


def test_estimate_radius_is_96_for_this_topology():
    p = O.init_params(1, 2, seed=3, base=8, dtype=torch.float64)
    r, g = O.estimate_radius(p, 1, seed=0)
    assert r == 96 and g.shape == (192, 192)


def test_batchnorm_fold_identities():
    """The algebra the next round's folded-BatchNorm kernels rely on: convolving the pre-BN activation with
    channel-scaled weights plus a 9-case border bias equals convolving the BatchNorm output, and the weight gradient follows
    from the gradient on `a` plus a border-sum correction -- exactly, in fp64, including 2-pixel images and negative scales."""
    from oracle import unet_numpy as ON
    rng = np.random.default_rng(0)
    for (N, H, W, Ci, Co) in [(2, 7, 9, 5, 4), (1, 2, 2, 3, 2), (3, 2, 6, 4, 3), (1, 16, 3, 2, 5)]:
        a = np.maximum(rng.normal(0.3, 1.0, size=(N, H, W, Ci)), 0)
        w = rng.normal(size=(3, 3, Ci, Co))
        b = rng.normal(size=Co)
        s = rng.normal(size=Ci)                      # gamma * rstd: either sign
        t = rng.normal(size=Ci)                      # beta - mean * s
        y = a * s + t
        ref = ON.conv_fwd(y, w, b)
        got = ON.conv_fwd_folded(a, w, b, s, t)
        assert np.abs(got - ref).max() < 1e-12 * max(1.0, np.abs(ref).max())
        dz = rng.normal(size=(N, H, W, Co))
        dw_ref, _ = ON.conv_wgrad(y, dz, 3)
        dw = ON.conv_wgrad_folded(a, dz, s, t)
        assert np.abs(dw - dw_ref).max() < 1e-11 * max(1.0, np.abs(dw_ref).max())
    # the interior bias is the plain bias plus the full tap sum; a corner drops 5 of the 9 taps
    Tt = rng.normal(size=(3, 3, 2))
    bias = ON.border_case_bias(Tt, np.zeros(2))
    assert np.allclose(bias[1, 1], Tt.sum((0, 1))) and np.allclose(bias[0, 0], Tt[1:, 1:].sum((0, 1))) and np.allclose(bias[2, 1], Tt[:2].sum((0, 1)))


def test_bn_backward_sums_from_the_consumer_weight_gradient():
    """the identity behind ub_bn_bwd_sums_wgrad: dbeta / dgamma of the producer's BatchNorm from the consumer's weight gradient on
    `a` and its border sums == the sums over the gradient tensor itself (fp64), incl. 2-pixel images and a concat consumer"""
    from oracle import unet_numpy as ON
    rng = np.random.default_rng(5)
    for (N, H, W, Cin, Cout) in [(2, 7, 9, 5, 4), (1, 2, 2, 3, 6), (3, 16, 4, 8, 2)]:
        a = np.maximum(rng.normal(0.3, 1, size=(N, H, W, Cin)), 0)
        dz = rng.normal(size=(N, H, W, Cout))
        w = rng.normal(size=(3, 3, Cin, Cout))
        mu, rstd = a.mean((0, 1, 2)), 1 / np.sqrt(a.var((0, 1, 2)) + 1e-3)
        dy = ON.conv_dgrad(dz, w)
        dw_a, _ = ON.conv_wgrad(a, dz, 3)
        dbeta, dgamma = ON.bn_bwd_sums_from_wgrad(w, dw_a, ON.border_sums(dz), mu, rstd)
        assert np.allclose(dbeta, dy.sum((0, 1, 2)), rtol=1e-10, atol=1e-10)
        assert np.allclose(dgamma, (dy * (a - mu) * rstd).sum((0, 1, 2)), rtol=1e-10, atol=1e-10)
