"""Pins the CPU oracle (oracle/dopt_ref.py) against every known-answer vector the reference's own unit tests hold for the
hot path, plus float64 closed forms / finite differences for the ops the reference never tests."""
import numpy as np
import pytest

from oracle import dopt_ref as R

F = np.float32


def f32(shape, vals):
    return np.array(vals, dtype=F).reshape(shape)


# ---- core/source/dopt/core/ops/nnet.d:270-293 ------------------------------------------------------------------------
def test_convolution_kat_is_true_convolution():
    x = f32([1, 1, 3, 5], [1, 1, 1, 0, 0] * 3)
    w = f32([1, 1, 1, 2], [-1, 1])
    y = R.convolution(x, w)
    assert y.shape == (1, 1, 3, 4)
    assert y.ravel().tolist() == [0, 0, 1, 0] * 3


# ---- core/source/dopt/core/ops/nnet.d:333-352 ------------------------------------------------------------------------
def test_maxpool_kat():
    x = f32([1, 1, 4, 4], [1, 2, 4, 3, 5, 3, 2, 2, 0.1, -4, 3, 2, 0, 0, 2, 2])
    assert R.maxpool(x, [2, 2]).ravel().tolist() == [F(5), F(4), F(0.1), F(3)]


# ---- core/source/dopt/core/ops/nnet.d:424-435 ------------------------------------------------------------------------
def test_softmax_kat():
    y = R.softmax(f32([1, 5], [1, 2, 3, 1, 2]))
    np.testing.assert_allclose(y.ravel(), [0.0674508, 0.18335, 0.498398, 0.0674508, 0.18335], rtol=1e-2, atol=1e-5)
    # D's approxEqual defaults: maxRelDiff 1e-2, maxAbsDiff 1e-5; the values actually agree to 1e-6
    np.testing.assert_allclose(y.ravel(), [0.0674508, 0.18335, 0.498398, 0.0674508, 0.18335], atol=1e-6)


# ---- core/source/dopt/core/ops/math.d:212-231, 285-298, 318-331, 361-373 ---------------------------------------------
def test_matmul_kat():
    c = R.matmul(f32([2, 1], [1, 2]), f32([1, 2], [3, 4]))
    assert c.ravel().tolist() == [3, 4, 6, 8]


def test_sum_kats():
    assert R.sum_(f32([2], [0.5, 1.5]), [0]).ravel().tolist() == [2.0]
    m = f32([2, 2], [0, 1, 2, 5])
    assert R.sum_(m, [0, 1]).ravel().tolist() == [8.0]
    # the rank-2 single-axis cases are lowered to matmul with a ones vector (math.d:257-279)
    assert R.matmul(np.ones((1, 2), F), m).ravel().tolist() == [2.0, 6.0]
    assert R.matmul(m, np.ones((2, 1), F)).ravel().tolist() == [1.0, 7.0]
    assert R.sum_(m, [0]).tolist() == [2.0, 6.0]
    assert R.sum_(m, [1]).tolist() == [1.0, 7.0]


def test_argmin_kats():
    assert R.argmin(f32([5], [4, 2, 6, 1, 2]), 0).ravel().tolist() == [3]
    assert R.argmin(f32([2, 3], [5, 1, 3, 6, 7, 2]), 1).ravel().tolist() == [1, 2]


def test_max_element_kats():
    a = f32([2, 2], [1, 4, 3, 6])
    assert R.max_element(a, [0, 1]).ravel().tolist() == [6.0]
    assert R.max_element(a, [0]).tolist() == [3.0, 6.0]
    assert R.max_element(a, [1]).tolist() == [4.0, 6.0]


# ---- core/source/dopt/core/ops/basic.d:212-395 -----------------------------------------------------------------------
def test_slice_pad_reshape_transpose_repeat_kats():
    s = R.slice_(np.arange(1, 10, dtype=np.int32).reshape(3, 3), [1, 1], [3, 3])
    assert s.ravel().tolist() == [5, 6, 8, 9]
    p = R.pad(np.array([[3]], dtype=np.int32), [2, 1], [3, 3])
    assert p.shape == (6, 5) and p[2, 1] == 3 and p.sum() == 3
    r = R.evaluate_op("reshape", [f32([2, 2], [1, 2, 3, 4])], {"shape": [1, 4]})
    assert r.shape == (1, 4) and r.ravel().tolist() == [1, 2, 3, 4]
    t = R.transpose(f32([2, 2], [1, 2, 3, 4]), [1, 0])
    assert t.ravel().tolist() == [1, 3, 2, 4]
    r1 = R.repeat(f32([1, 1], [3]), [2, 3])
    assert r1.ravel().tolist() == [3] * 6
    r2 = R.repeat(f32([2, 2], [1, 2, 3, 4]), [3, 2])
    assert r2.ravel().tolist() == [1, 2, 1, 2, 3, 4, 3, 4] * 3
    # repeat(n) is lowered to matmul(ones[n,1], vec[1,V]) (basic.d:370-381)
    r3 = R.matmul(np.ones((3, 1), F), f32([1, 2], [1, 2]))
    assert r3.ravel().tolist() == [1, 2, 1, 2, 1, 2]


# ---- cuda/source/dopt/cuda/package.d:533-542 -------------------------------------------------------------------------
def test_cuda_smoke_kat():
    a, b, c = f32([], [3]), f32([], [4]), f32([], [-1])
    assert R.POINTWISE_BINARY["add"](R.POINTWISE_BINARY["mul"](a, b), c) == F(11)


# ---- core/source/dopt/core/grads/basic.d:49-67 (sliceGrad -> pad) -----------------------------------------------------
def test_slice_grad_kat():
    b = f32([4, 4], [5, 6, 7, 8] * 4)
    pg = np.ones((1, 1), F)
    g = R.pad(pg, [1, 1], [2, 2]) * b  # d/da of slice(a*b,[1,1],[2,2])
    assert g.ravel().tolist() == [0, 0, 0, 0, 0, 6, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]


# ---- ops without a reference test: closed forms / finite differences in float64 -------------------------------------
def _fd(fn, x, eps=1e-3):
    g = np.zeros_like(x, dtype=np.float64)
    it = np.nditer(x, flags=["multi_index"])
    while not it.finished:
        i = it.multi_index
        xp, xm = x.copy(), x.copy()
        xp[i] += eps
        xm[i] -= eps
        g[i] = (fn(xp) - fn(xm)) / (2 * eps)
        it.iternext()
    return g


@pytest.mark.parametrize("pad,stride", [((0, 0), (1, 1)), ((1, 1), (1, 1)), ((1, 1), (2, 2)), ((2, 1), (1, 2))])
def test_conv_gradients_match_finite_differences(pad, stride):
    rng = np.random.RandomState(0)
    x = rng.randn(2, 3, 6, 7).astype(F)
    w = rng.randn(4, 3, 3, 2).astype(F)
    y = R.convolution(x, w, pad, stride)
    dy = rng.randn(*y.shape).astype(F)
    loss_x = lambda xx: float((R.convolution(xx.astype(F), w, pad, stride).astype(np.float64) * dy).sum())
    loss_w = lambda ww: float((R.convolution(x, ww.astype(F), pad, stride).astype(np.float64) * dy).sum())
    dx = R.convolution_features_grad(dy, w, x.shape, pad, stride)
    dw = R.convolution_filters_grad(dy, x, w.shape, pad, stride)
    np.testing.assert_allclose(dx, _fd(loss_x, x.astype(np.float64)), rtol=2e-2, atol=2e-2)
    np.testing.assert_allclose(dw, _fd(loss_w, w.astype(np.float64)), rtol=2e-2, atol=2e-2)


def test_conv_direct_loop_definition():
    """The im2col formulation equals the reference's own nested loops (cpu/source/dopt/cpu/nnet.d:30-60, with the input
    bound check fixed as documented in the oracle header)."""
    rng = np.random.RandomState(1)
    x = rng.randn(1, 2, 5, 5).astype(F)
    w = rng.randn(3, 2, 3, 3).astype(F)
    pad, stride = (1, 1), (2, 2)
    y = R.convolution(x, w, pad, stride)
    ref = np.zeros_like(y)
    for k in range(3):
        for c in range(2):
            for p in range(y.shape[2]):
                for q in range(y.shape[3]):
                    for j in range(3):
                        for i in range(3):
                            iy, ix = p * 2 - 1 + j, q * 2 - 1 + i
                            if 0 <= iy < 5 and 0 <= ix < 5:
                                ref[0, k, p, q] += w[k, c, 2 - j, 2 - i] * x[0, c, iy, ix]
    np.testing.assert_allclose(y, ref, rtol=1e-5, atol=1e-5)


def test_batchnorm_train_grad_inference():
    rng = np.random.RandomState(2)
    x = (rng.randn(4, 3, 5, 5) * 2 + 1).astype(F)
    g, b = rng.rand(3).astype(F) + 0.5, rng.randn(3).astype(F)
    rm, rv = rng.randn(3).astype(F), rng.rand(3).astype(F) + 0.5
    packed = R.batch_norm_train(x, g.reshape(1, 3, 1, 1), b, rm, rv, 0.9)
    V = x.size
    y = packed[:V].reshape(x.shape)
    np.testing.assert_allclose(y.mean(axis=(0, 2, 3)), b, atol=1e-5)
    np.testing.assert_allclose(y.std(axis=(0, 2, 3)), g, rtol=1e-3)
    M = 4 * 25
    np.testing.assert_allclose(packed[V:V + 3], 0.9 * rm + 0.1 * x.mean(axis=(0, 2, 3)), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(packed[V + 3:], 0.9 * rv + 0.1 * x.var(axis=(0, 2, 3)) * M / (M - 1), rtol=1e-5)
    dy = rng.randn(*x.shape).astype(F)

    def loss(xx):
        p = R.batch_norm_train(xx.astype(F), g, b, rm, rv, 0.9)[:V].astype(np.float64)
        return float((p * dy.ravel()).sum())
    pk = R.batch_norm_grad(dy, x, g, out_volume=2 * V + 3)
    assert pk.size == 2 * V + 3 and not pk[V + 6:].any()
    np.testing.assert_allclose(pk[:V].reshape(x.shape), _fd(loss, x.astype(np.float64), 1e-2), rtol=5e-2, atol=5e-3)
    xhat = (x - x.mean(axis=(0, 2, 3), keepdims=True)) / np.sqrt(x.var(axis=(0, 2, 3), keepdims=True) + 1e-5)
    np.testing.assert_allclose(pk[V:V + 3], (dy * xhat).sum(axis=(0, 2, 3)), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(pk[V + 3:V + 6], dy.sum(axis=(0, 2, 3)), rtol=1e-4, atol=1e-4)
    yi = R.batch_norm_inference(x, g, b, rm, rv)
    np.testing.assert_allclose(yi, g.reshape(1, 3, 1, 1) * (x - rm.reshape(1, 3, 1, 1)) /
                               np.sqrt(rv.reshape(1, 3, 1, 1) + 1e-5) + b.reshape(1, 3, 1, 1), rtol=1e-5, atol=1e-5)


def test_batchnorm_running_mean_kat():
    """nnet/source/dopt/nnet/layers/batchnorm.d:158-177: after many steps the running mean of x=[[1,2],[3,4],[5,6]]
    converges to the batch mean [3,4] (the projection replaces `mean` by the packed update every step)."""
    x = f32([3, 2], [1, 2, 3, 4, 5, 6])
    g, b = np.ones(2, F), np.zeros(2, F)
    rm, rv = np.zeros(2, F), np.ones(2, F)
    for _ in range(1000):
        packed = R.batch_norm_train(x, g, b, rm, rv, 0.9)
        rm, rv = packed[6:8], packed[8:10]
    np.testing.assert_allclose(rm, [3.0, 4.0], rtol=1e-2, atol=1e-5)


def test_relu_pool_softmax_bias_grads():
    rng = np.random.RandomState(3)
    x = rng.randn(2, 3, 4, 4).astype(F)
    dy = rng.randn(2, 3, 4, 4).astype(F)
    np.testing.assert_array_equal(R.relu(x), np.maximum(x, 0))
    np.testing.assert_array_equal(R.relu_grad(dy, R.relu(x), x), dy * (x > 0))
    y = R.maxpool(x, [2, 2])
    dyp = rng.randn(*y.shape).astype(F)
    dx = R.maxpool_grad(dyp, y, x, [2, 2])
    assert np.count_nonzero(dx) == y.size and np.isclose(dx.sum(), dyp.sum(), rtol=1e-5)
    # ties: all-zero window (common after ReLU)
    z = np.zeros((1, 1, 2, 2), F)
    assert R.maxpool_grad(np.ones((1, 1, 1, 1), F), R.maxpool(z, [2, 2]), z, [2, 2], tie_all=True).sum() == 4
    assert R.maxpool_grad(np.ones((1, 1, 1, 1), F), R.maxpool(z, [2, 2]), z, [2, 2], tie_all=False).ravel().tolist() == [1, 0, 0, 0]
    l = rng.randn(5, 7).astype(F)
    p = R.softmax(l)
    g = rng.randn(5, 7).astype(F)
    fd = _fd(lambda ll: float((R.softmax(ll.astype(F)).astype(np.float64) * g).sum()), l.astype(np.float64))
    np.testing.assert_allclose(R.softmax_grad(g, p), fd, rtol=2e-2, atol=2e-3)
    bias = rng.randn(3).astype(F)
    np.testing.assert_allclose(R.add_bias(x, bias), x + bias.reshape(1, 3, 1, 1))
    np.testing.assert_allclose(R.add_bias_grad(dy), dy.sum(axis=(0, 2, 3)), rtol=1e-5, atol=1e-5)


def test_pointwise_semantics():
    a = np.array([-2.5, 0.0, 3.0], F)
    b = np.array([1.0, 0.0, -4.0], F)
    assert R.POINTWISE_UNARY["sgn"](a).tolist() == [-1, 0, 1]
    assert R.POINTWISE_BINARY["lt"](a, b).tolist() == [1, 0, 0]
    assert R.POINTWISE_BINARY["eq"](a, b).tolist() == [0, 1, 0]
    ai = np.array([7, -7, 9], np.int32)
    bi = np.array([2, 2, 0], np.int32)
    assert R.POINTWISE_BINARY["div"](ai, bi).tolist() == [3, -3, 0]
    assert R.POINTWISE_UNARY["sqrt"](np.array([10], np.int32)).tolist() == [3]
    assert R.POINTWISE_UNARY["exp"](np.array([2], np.int32)).tolist() == [7]


def test_optimiser_steps_follow_the_reference_graphs():
    rng = np.random.RandomState(4)
    w, g, m = (rng.randn(16).astype(F) for _ in range(3))
    nw, nm = R.sgd_step(w, g, m, 0.1, 0.9)
    np.testing.assert_array_equal(nm, (m * F(0.9) + F(0.1) * g).astype(F))
    np.testing.assert_array_equal(nw, w - nm)
    v = np.abs(rng.randn(16)).astype(F)
    out = R.adam_step(w, g, m, v, 1.0, 1.0, 1e-3, 0.9, 0.999, 1e-8, vhat=np.zeros(16, F))
    assert np.isclose(out[3], 0.9) and np.isclose(out[4], 0.999)
    np.testing.assert_array_equal(out[5], np.maximum(0, v))   # max(varhat, OLD var): survey F11
