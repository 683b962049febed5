"""Pins the numpy oracle (oracle/dopt_ref.py) against outputs of the reference CUDA backend's own library calls: the
vectors in tests/golden/cudnn_golden.npz were produced on a B200 by tools/make_cudnn_golden.py, which pushes seeded inputs
through the cuDNN / cuBLAS call sequences of cuda/source/dopt/cuda/nnet/cudnn7.d, math.d and basic.d (replayed from C++,
oracle/cudnn_replay.cpp).  Runs on CPU: this is how the ops with no reference unit test (conv gradients, maxpoolGrad and
its tie routing, softmaxGrad, relu(+Grad), addBias(+Grad), batchNorm train / grad / inference) stop being "parity unpinned".

Tolerance: bit-exact for relu / pooling / addBias / transpose; 2e-6 of the tensor's max magnitude for everything that sums
(cuDNN's summation order differs from numpy's)."""
import os

import numpy as np
import pytest

from oracle import dopt_ref as R

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cudnn_golden.npz"))
TOL = 2e-6


def g(case, name):
    return G["%s/%s" % (case, name)]


def close(got, want, tol=TOL):
    scale = max(float(np.abs(want).max()), 1e-30)
    err = float(np.abs(got.astype(np.float64) - want.astype(np.float64)).max()) / scale
    assert got.shape == want.shape and err < tol, err


def test_fixture_comes_from_cudnn_9():
    assert int(G["meta/cudnn"]) >= 90000


@pytest.mark.parametrize("case", ["conv0", "conv1", "conv2", "conv3"])
def test_convolution_family(case):
    x, w, dy = g(case, "x"), g(case, "w"), g(case, "dy")
    p, s = g(case, "padding").tolist(), g(case, "stride").tolist()
    close(R.convolution(x, w, p, s), g(case, "y"))
    close(R.convolution_features_grad(dy, w, x.shape, p, s), g(case, "dx"))
    close(R.convolution_filters_grad(dy, x, w.shape, p, s), g(case, "dw"))


@pytest.mark.parametrize("case", ["pool0", "pool1"])
def test_maxpool_and_tie_routing(case):
    x, dims = g(case, "x"), g(case, "dims").tolist()
    y = R.maxpool(x, dims)
    np.testing.assert_array_equal(y, g(case, "y"))
    dx = R.maxpool_grad(g(case, "dy"), y, x, dims)
    np.testing.assert_array_equal(dx, g(case, "dx"))
    # the fixture contains tied windows (zeros after a relu, and one constant window): cuDNN gave the whole gradient to
    # the first maximum, so routing to every tied element must NOT reproduce it
    assert not np.array_equal(R.maxpool_grad(g(case, "dy"), y, x, dims, tie_all=True), g(case, "dx"))


@pytest.mark.parametrize("case", ["softmax0", "softmax1"])
def test_softmax_and_grad(case):
    y = R.softmax(g(case, "x"))
    close(y, g(case, "y"))
    close(R.softmax_grad(g(case, "dy"), g(case, "y")), g(case, "dx"))


def test_relu_and_grad():
    x = g("relu", "x")
    np.testing.assert_array_equal(R.relu(x), g("relu", "y"))
    np.testing.assert_array_equal(R.relu_grad(g("relu", "dy"), g("relu", "y"), x), g("relu", "dx"))


def test_add_bias_and_grad():
    x, b = g("bias", "x"), g("bias", "b")
    np.testing.assert_array_equal(R.add_bias(x, b), g("bias", "y"))
    close(R.add_bias_grad(x), g("bias", "db"))
    # the reference accumulates into a buffer zeroed once -- its second execution returns twice the gradient
    close(2 * g("bias", "db"), g("bias", "db_second_execution"))


@pytest.mark.parametrize("case", ["bn0", "bn1"])
def test_batchnorm_train_grad_inference(case):
    x, sc, b, rm, rv, dy = (g(case, k) for k in ("x", "scale", "bias", "mean", "var", "dy"))
    V, C = x.size, x.shape[1]
    got = R.batch_norm_train(x, sc, b, rm, rv, float(g(case, "momentum")))
    want = g(case, "train")
    close(got[:V], want[:V])
    close(got[V:V + C], want[V:V + C])                     # running mean
    close(got[V + C:], want[V + C:])                       # running variance (unbiased batch variance blended in)
    gg = R.batch_norm_grad(dy, x, sc, 2 * V + C)
    wg = g(case, "grad")
    close(gg[:V], wg[:V], 1e-5)
    close(gg[V:V + C], wg[V:V + C], 1e-5)                  # dscale
    close(gg[V + C:V + 2 * C], wg[V + C:V + 2 * C], 1e-5)  # dbias
    assert not gg[V + 2 * C:].any() and not wg[V + 2 * C:].any()
    close(R.batch_norm_inference(x, sc, b, rm, rv), g(case, "inference"))


def test_matmul_and_transpose():
    close(R.matmul(g("matmul", "a"), g("matmul", "b")), g("matmul", "c"))
    np.testing.assert_array_equal(R.transpose(g("transpose", "a"), [1, 0]), g("transpose", "c"))
