"""Property tests (hypothesis, CPU): the host mirror's shape rules (dopt's `judge` functions) against the oracle's results on
random geometries -- convolution (core/ops/nnet.d:43-87), convolutionTranspose (nnet.d:305-315), maxpool (nnet.d:89-107),
slice / pad / transpose / repeat (core/ops/basic.d) -- and the gradient graphs of those ops against each other
(conv dgrad of a conv output has the input's shape; pad undoes slice)."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from dopt_b200 import host as H
from oracle import dopt_ref as R
from oracle import graph_eval as G

F = np.float32
FAST = settings(max_examples=40, deadline=None)


@pytest.fixture(autouse=True)
def _fresh():
    H.init()
    H.reset()
    yield
    H.reset()


@FAST
@given(n=st.integers(1, 3), c=st.integers(1, 4), k=st.integers(1, 4), h=st.integers(3, 9), w=st.integers(3, 9),
       r=st.integers(1, 3), s=st.integers(1, 3), pad=st.integers(0, 2), stride=st.integers(1, 3), seed=st.integers(0, 999))
def test_convolution_judge_and_transpose_shapes(n, c, k, h, w, r, s, pad, stride, seed):
    H.reset()
    rng = np.random.RandomState(seed)
    xv, wv = rng.randn(n, c, h, w).astype(F), rng.randn(k, c, r, s).astype(F)
    x, f = H.float32((n, c, h, w), xv), H.float32((k, c, r, s), wv)
    y = H.convolution(x, f, (pad, pad), (stride, stride))
    ref = R.convolution(xv, wv, (pad, pad), (stride, stride))
    assert y.shape == ref.shape == (n, k, (h + 2 * pad - r) // stride + 1, (w + 2 * pad - s) // stride + 1)
    got = G.evaluate_ops(H, [y])[0]
    np.testing.assert_array_equal(got, ref)
    # convolutionTranspose of the result has the shape the reference's formula gives: (P-1)*u + R - 2*pad
    t = H.convolution_transpose(y, f, (pad, pad), (stride, stride))
    assert t.shape == (n, c, (y.shape[2] - 1) * stride + r - 2 * pad, (y.shape[3] - 1) * stride + s - 2 * pad)
    # the gradient graph w.r.t. the features has the features' shape and equals the oracle's feature gradient of ones
    g = H.grad(H.sum_(y), [x])[0]
    assert g.shape == (n, c, h, w)
    np.testing.assert_allclose(G.evaluate_ops(H, [g])[0],
                               R.convolution_features_grad(np.ones_like(ref), wv, xv.shape, (pad, pad), (stride, stride)),
                               rtol=1e-5, atol=1e-6)


@FAST
@given(n=st.integers(1, 3), c=st.integers(1, 3), h=st.integers(2, 9), w=st.integers(2, 9), dh=st.integers(1, 3),
       dw=st.integers(1, 3), seed=st.integers(0, 999))
def test_maxpool_floor_shape_and_gradient_mass(n, c, h, w, dh, dw, seed):
    if h < dh or w < dw:
        return
    H.reset()
    xv = np.random.RandomState(seed).randn(n, c, h, w).astype(F)
    x = H.float32((n, c, h, w), xv)
    y = H.maxpool(x, (dh, dw))
    assert y.shape == (n, c, h // dh, w // dw)
    g = H.grad(H.sum_(y), [x])[0]
    gv = G.evaluate_ops(H, [g])[0]
    assert gv.shape == xv.shape
    assert gv.sum() == y.shape[2] * y.shape[3] * n * c          # one unit of gradient per window, ties or not
    assert not gv[:, :, (h // dh) * dh:, :].any() and not gv[:, :, :, (w // dw) * dw:].any()


@FAST
@given(data=st.data())
def test_slice_pad_transpose_repeat_against_numpy(data):
    H.reset()
    rank = data.draw(st.integers(1, 3))
    shape = tuple(data.draw(st.integers(1, 5)) for _ in range(rank))
    xv = np.arange(int(np.prod(shape)), dtype=F).reshape(shape)
    x = H.float32(shape, xv)
    start = [data.draw(st.integers(0, d - 1)) for d in shape]
    stop = [data.draw(st.integers(s + 1, d)) for s, d in zip(start, shape)]
    sl = H.slice_(x, start, stop)
    want = xv[tuple(slice(a, b) for a, b in zip(start, stop))]
    np.testing.assert_array_equal(G.evaluate_ops(H, [sl])[0], want)
    # pad with the slice's complement restores the extent (this is sliceGrad, grads/basic.d:20-31)
    back = H.pad(sl, start, [d - b for d, b in zip(shape, stop)])
    assert back.shape == shape
    restored = G.evaluate_ops(H, [back])[0]
    np.testing.assert_array_equal(restored[tuple(slice(a, b) for a, b in zip(start, stop))], want)
    assert restored.sum() == want.sum()
    reps = [data.draw(st.integers(1, 3)) for _ in range(rank)]
    np.testing.assert_array_equal(G.evaluate_ops(H, [H.repeat(x, reps)])[0], np.tile(xv, reps))
    if rank == 2:
        np.testing.assert_array_equal(G.evaluate_ops(H, [H.transpose(x, [1, 0])])[0], xv.T)
