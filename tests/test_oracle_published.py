"""Known-answer vectors for the oracle's building blocks taken from the PUBLISHED definitions of the third-party ops the
reference calls (TensorFlow / Keras API documentation examples and formulas; TensorFlow itself cannot be installed here).
They pin the oracle at the level of single ops -- loss, optimiser, pooling, softmax, dropout scaling, BatchNorm algebra,
convolution orientation -- not at the level of the whole graph, which stays unpinned."""
import math

import numpy as np
import torch

from oracle import unet_oracle as O


def test_categorical_crossentropy_documentation_example():
    # tf.keras.losses.CategoricalCrossentropy docs: y_true [[0,1,0],[0,0,1]], y_pred [[0.05,0.95,0],[0.1,0.8,0.1]]
    # -> reduction NONE [0.0513, 2.303], mean 1.177, SUM 2.354.  The reference uses reduction NONE and then
    # reduce_sum(axis=0) / global_batch (UNet/model.py:77, :211-215).  Keras clips probabilities to [1e-7, 1 - 1e-7].
    y_true = np.array([[0, 1, 0], [0, 0, 1]])
    y_pred = np.clip(np.array([[0.05, 0.95, 0.0], [0.1, 0.8, 0.1]]), 1e-7, 1 - 1e-7)
    logits = torch.tensor(np.log(y_pred)).reshape(2, 1, 1, 3)                 # softmax(log p) == p for normalised p
    onehot = torch.tensor(y_true).reshape(2, 1, 1, 3)
    ce = -(onehot.double() * torch.log_softmax(logits, dim=-1)).sum(-1).reshape(2)
    assert np.allclose(ce.numpy(), [0.0513, 2.303], atol=5e-4)
    loss, acc = O.loss_and_accuracy(logits, onehot, 2)                        # sum over the batch / global batch, mean over pixels
    assert abs(float(loss) - 1.177) < 5e-4 and float(acc) == 0.5
    loss4, _ = O.loss_and_accuracy(logits, onehot, 4)                         # a replica holding half of a global batch of 4
    assert abs(float(loss4) - 2.354 / 4) < 5e-4


def test_adam_documentation_example():
    # tf.keras.optimizers.Adam docs: learning_rate 0.1, var = 10.0, loss = var**2 / 2 (gradient == var): after one step the
    # variable is 9.9 ("the first step is -learning_rate * sign(grad)")
    p = {"w/kernel": torch.tensor([10.0], dtype=torch.float64)}
    opt = O.KerasAdam.__new__(O.KerasAdam)
    opt.lr, opt.t = 0.1, 0
    opt.m = {"w/kernel": torch.zeros(1, dtype=torch.float64)}
    opt.v = {"w/kernel": torch.zeros(1, dtype=torch.float64)}
    opt.apply(p, {"w/kernel": p["w/kernel"].clone()})
    assert abs(float(p["w/kernel"]) - 9.9) < 1e-6
    # epsilon sits outside the bias correction (Keras: lr_t * m / (sqrt(v) + eps)): a tiny gradient shows the difference to the
    # textbook form m_hat / (sqrt(v_hat) + eps)
    g = 1e-8
    lr_t = 0.1 * math.sqrt(1 - 0.999) / (1 - 0.9)
    keras = lr_t * (0.1 * g) / (math.sqrt(0.001 * g * g) + 1e-7)
    textbook = 0.1 * g / (g + 1e-7)
    q = {"w/kernel": torch.tensor([0.0], dtype=torch.float64)}
    opt.t = 0
    opt.m["w/kernel"].zero_()
    opt.v["w/kernel"].zero_()
    opt.apply(q, {"w/kernel": torch.tensor([g], dtype=torch.float64)})
    assert abs(-float(q["w/kernel"]) - keras) < 1e-15 and abs(keras - textbook) > 0.1 * textbook


def test_maxpool_softmax_dropout_documentation_examples():
    # tf.keras.layers.MaxPool2D docs: [[1,2,3,4],[5,6,7,8],[9,10,11,12]], pool 2x2, strides 2, 'valid' -> [[6, 8]]
    x = torch.tensor([[1., 2., 3., 4.], [5., 6., 7., 8.], [9., 10., 11., 12.]]).reshape(1, 1, 3, 4)
    assert O._pool(x[:, :, :2]).reshape(-1).tolist() == [6.0, 8.0]
    # tf.keras.layers.Softmax docs: [1., 2., 1.] -> [0.21194157, 0.5761169, 0.21194157]
    sm = torch.softmax(torch.tensor([1.0, 2.0, 1.0], dtype=torch.float64), dim=-1)
    assert np.allclose(sm.numpy(), [0.21194157, 0.5761169, 0.21194157], atol=1e-7)
    # tf.keras.layers.Dropout docs: kept inputs are scaled by 1 / (1 - rate); rate 0.5 -> x2 (UNet/model.py:62)
    keep = torch.tensor([[1, 0], [0, 1]]).reshape(1, 1, 2, 2)
    out = O._dropout(torch.ones(1, 1, 2, 2), keep, True)
    assert out.reshape(-1).tolist() == [2.0, 0.0, 0.0, 2.0] and torch.equal(O._dropout(torch.ones(1, 1, 2, 2), keep, False), torch.ones(1, 1, 2, 2))


def test_batchnorm_published_formula():
    # tf.keras.layers.BatchNormalization docs: training -> gamma * (batch - mean(batch)) / sqrt(var(batch) + epsilon) + beta with
    # the (biased) batch variance; moving = moving * momentum + batch * (1 - momentum); defaults momentum 0.99, epsilon 1e-3
    rng = np.random.default_rng(0)
    x = torch.tensor(rng.normal(2.0, 3.0, size=(4, 3, 5, 6)))
    p = {"l/gamma": torch.tensor([1.5, 0.5, -1.0], dtype=torch.float64), "l/beta": torch.tensor([0.1, -0.2, 0.3], dtype=torch.float64),
         "l/moving_mean": torch.zeros(3, dtype=torch.float64), "l/moving_var": torch.ones(3, dtype=torch.float64)}
    new = {}
    y = O._bn(x, "l", p, True, new)
    xn = x.numpy()
    mean, var = xn.mean((0, 2, 3)), xn.var((0, 2, 3))
    ref = p["l/gamma"].numpy()[None, :, None, None] * (xn - mean[None, :, None, None]) / np.sqrt(var + 1e-3)[None, :, None, None] + p["l/beta"].numpy()[None, :, None, None]
    assert np.abs(y.numpy() - ref).max() < 1e-12
    assert np.allclose(new["l/moving_mean"].numpy(), 0.01 * mean)
    n = 4 * 5 * 6
    assert np.allclose(new["l/moving_var"].numpy(), 0.99 + 0.01 * var * n / (n - 1))          # fused kernel: unbiased variance
    # inference: the moving statistics
    y2 = O._bn(x, "l", p, False, None)
    ref2 = p["l/gamma"].numpy()[None, :, None, None] * xn / np.sqrt(1.0 + 1e-3) + p["l/beta"].numpy()[None, :, None, None]
    assert np.abs(y2.numpy() - ref2).max() < 1e-12


def test_convolution_orientation_and_transpose_layout():
    # Conv2D is a cross-correlation with kernel [kh, kw, Cin, Cout] and 'same' zero padding; Conv2DTranspose(2, strides 2) with
    # kernel [kh, kw, Cout, Cin] writes out[2i+a, 2j+b, co] = bias + sum_ci in[i, j, ci] * W[a, b, co, ci]      
    from oracle import unet_numpy as ON
    x = np.zeros((1, 3, 3, 1))
    x[0, 1, 1, 0] = 1.0                                          # a delta at the centre
    w = np.arange(9, dtype=np.float64).reshape(3, 3, 1, 1)
    out = ON.conv_fwd(x, w, np.zeros(1))[0, :, :, 0]
    assert np.array_equal(out, w[::-1, ::-1, 0, 0])             # cross-correlation: the response to a delta is the FLIPPED kernel
    xd = np.zeros((1, 1, 2, 2))
    xd[0, 0, 0, 1] = 3.0                                         # [N, h, w, Cin] with Cin = 2: pixel (0,0), channel 1
    wt = np.arange(16, dtype=np.float64).reshape(2, 2, 2, 2)    # [a, b, Cout, Cin]
    od = ON.deconv_fwd(xd, wt, np.array([0.5, -0.5]))
    assert od.shape == (1, 2, 4, 2)
    for a in range(2):
        for b in range(2):
            assert np.allclose(od[0, a, b], 3.0 * wt[a, b, :, 1] + np.array([0.5, -0.5]))
    assert np.allclose(od[0, :, 2:], np.array([0.5, -0.5]))     # the other input pixel is zero: bias only
