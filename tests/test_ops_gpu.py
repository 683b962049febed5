"""Per-op parity on a real B200: every kernel dopt's CUDA backend registers, called through the C ABI exactly as
CUDAKernel.execute would be, against the CPU oracle on identical seeded inputs.

Tolerances (stated per north star):
  * integer / copy / comparison / add / sub / mul / div / sqrt / relu / pooling results: bit-exact
  * exp / log / pow: 2 ulp-ish, rtol 2e-6 (CUDA libm vs numpy libm)
  * fp32 reductions (sum, softmax, batch-norm, bias grad), fp32 convolution / matmul (MATH_FP32): rtol 1e-4 of the
    tensor's max magnitude (summation order differs from the serial CPU loop)
  * bf16 tensor-core convolution / matmul (MATH_BF16): 2e-2 of the tensor's max magnitude (operands rounded to bf16,
    fp32 accumulation)
"""
import numpy as np
import pytest
import torch

import dopt_b200 as db
from oracle import dopt_ref as R

pytestmark = pytest.mark.gpu
F = np.float32


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    torch.cuda.synchronize()
    return t.cpu().numpy()


def run(op, inputs, attrs=None, math=db.MATH_DEFAULT, out_shape=None):
    attrs = attrs or {}
    ref = R.evaluate_op(op, inputs, attrs, out_shape)
    shape = out_shape if out_shape is not None else ref.shape
    odt = db.INT32 if ref.dtype == np.int32 else db.FLOAT32
    out = db.run_op(op, [dev(a) for a in inputs], shape, odt, math, **attrs)
    return host(out), ref


def rel_err(got, ref):
    scale = max(float(np.abs(ref).max()), 1e-30)
    return float(np.abs(got.astype(np.float64) - ref.astype(np.float64)).max()) / scale


# ---------------------------------------------------------------------------------------------------------------------
# pointwise
# ---------------------------------------------------------------------------------------------------------------------
EXACT_BIN = ["add", "sub", "mul", "div", "lt", "lte", "gt", "gte", "eq", "neq", "max", "min"]


@pytest.mark.parametrize("n", [1, 3, 4, 1023, 4096 + 5, 1 << 20])
@pytest.mark.parametrize("op", EXACT_BIN)
def test_pointwise_binary_float_exact(op, n):
    rng = np.random.RandomState(hash(op) % 1000 + n % 97)
    a = rng.randn(n).astype(F)
    b = rng.randn(n).astype(F)
    b[::7] = a[::7]  # make eq / lte / gte interesting
    got, ref = run(op, [a, b])
    np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize("op", EXACT_BIN + ["pow"])
def test_pointwise_binary_int(op):
    rng = np.random.RandomState(5)
    a = rng.randint(-9, 10, size=1001).astype(np.int32)
    b = rng.randint(1, 5, size=1001).astype(np.int32)
    got, ref = run(op, [a, b])
    np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize("op", ["neg", "abs", "sgn", "sqrt"])
def test_pointwise_unary_exact(op):
    rng = np.random.RandomState(6)
    a = rng.randn(100003).astype(F)
    if op == "sqrt":
        a = np.abs(a)
    got, ref = run(op, [a])
    np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize("op", ["exp", "log"])
def test_pointwise_unary_transcendental(op):
    rng = np.random.RandomState(7)
    a = rng.rand(100003).astype(F) * 4 + 0.01
    got, ref = run(op, [a])
    np.testing.assert_allclose(got, ref, rtol=2e-6, atol=1e-7)


# the 12 unary functions the reference's CUDA backend evaluates on the host (cuda/source/dopt/cuda/package.d:81-119,
# cpu/source/dopt/cpu/math.d:323-324): CUDA's single-precision functions are within 2 ulp of the correctly rounded value
@pytest.mark.parametrize("op", ["sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh", "asinh", "acosh", "atanh"])
def test_pointwise_unary_trigonometric(op):
    rng = np.random.RandomState(9)
    a = (rng.rand(100003).astype(F) * 2 - 1) * F(0.98)          # (-0.98, 0.98): inside every domain but acosh's
    if op == "acosh":
        a = np.abs(a) * 5 + F(1.01)
    elif op in ("sin", "cos", "tan", "atan", "sinh", "cosh", "tanh", "asinh"):
        a = a * F(1.5)                                          # tan stays away from its poles
    got, ref = run(op, [a])
    np.testing.assert_allclose(got, ref, rtol=4e-6, atol=2e-7)
    ai = np.arange(-3, 4, dtype=np.int32) if op != "acosh" else np.arange(1, 8, dtype=np.int32)
    if op in ("asin", "acos", "atanh"):
        ai = np.array([0, 0, 0], np.int32)
    got, ref = run(op, [ai])
    np.testing.assert_array_equal(got, ref)


def test_pointwise_pow_float():
    rng = np.random.RandomState(8)
    a = rng.rand(5000).astype(F) * 3 + 0.1
    b = rng.randn(5000).astype(F)
    got, ref = run("pow", [a, b])
    np.testing.assert_allclose(got, ref, rtol=5e-6)


def test_pointwise_unary_int():
    a = np.arange(0, 50, dtype=np.int32)
    for op in ["neg", "abs", "sgn", "sqrt"]:
        got, ref = run(op, [a - 10 if op != "sqrt" else a])
        np.testing.assert_array_equal(got, ref)


def test_reference_cuda_smoke_kat():
    # cuda/source/dopt/cuda/package.d:533-542: 3*4 + (-1) == 11
    m, _ = run("mul", [np.array(3, F).reshape(()), np.array(4, F).reshape(())])
    s, _ = run("add", [m, np.array(-1, F).reshape(())])
    assert s == F(11)


# ---------------------------------------------------------------------------------------------------------------------
# slice / pad / repeat / transpose (bit-exact)
# ---------------------------------------------------------------------------------------------------------------------
def test_basic_reference_kats():
    got, _ = run("slice", [np.arange(1, 10, dtype=np.int32).reshape(3, 3)], {"start": [1, 1], "stop": [3, 3]})
    assert got.ravel().tolist() == [5, 6, 8, 9]
    got, _ = run("pad", [np.array([[3]], np.int32)], {"before": [2, 1], "after": [3, 3]})
    assert got.shape == (6, 5) and got[2, 1] == 3 and got.sum() == 3
    got, _ = run("transpose", [np.array([[1, 2], [3, 4]], F)], {"order": [1, 0]})
    assert got.ravel().tolist() == [1, 3, 2, 4]
    got, _ = run("repeat", [np.array([[3]], F)], {"repetitions": [2, 3]})
    assert got.ravel().tolist() == [3] * 6
    got, _ = run("repeat", [np.array([[1, 2], [3, 4]], F)], {"repetitions": [3, 2]})
    assert got.ravel().tolist() == [1, 2, 1, 2, 3, 4, 3, 4] * 3


@pytest.mark.parametrize("shape,start,stop", [
    ((1000,), (10,), (990,)),                  # rank-1: the batch-norm unpack case
    ((37, 53), (3, 5), (30, 50)),
    ((4, 5, 6, 7), (1, 0, 2, 3), (3, 5, 5, 7)),
    ((8, 16, 8, 8), (2, 0, 0, 0), (5, 16, 8, 8)),   # contiguous outer cut
])
def test_slice_and_its_gradient_pad(shape, start, stop):
    rng = np.random.RandomState(9)
    x = rng.randn(*shape).astype(F)
    got, ref = run("slice", [x], {"start": list(start), "stop": list(stop)})
    np.testing.assert_array_equal(got, ref)
    after = [s - e for s, e in zip(shape, stop)]
    got2, ref2 = run("pad", [ref], {"before": list(start), "after": after})
    np.testing.assert_array_equal(got2, ref2)
    assert got2.shape == shape


@pytest.mark.parametrize("rows,cols", [(1, 1), (100, 640), (640, 100), (33, 65), (1, 77)])
def test_transpose(rows, cols):
    x = np.random.RandomState(10).randn(rows, cols).astype(F)
    got, ref = run("transpose", [x], {"order": [1, 0]})
    np.testing.assert_array_equal(got, ref)


def test_repeat_rank3():
    x = np.random.RandomState(11).randn(3, 1, 5).astype(F)
    got, ref = run("repeat", [x], {"repetitions": [2, 4, 3]})
    np.testing.assert_array_equal(got, ref)


# ---------------------------------------------------------------------------------------------------------------------
# reductions
# ---------------------------------------------------------------------------------------------------------------------
def test_reduction_reference_kats():
    got, _ = run("sum", [np.array([0.5, 1.5], F)], {"axes": [0]})
    assert got.ravel().tolist() == [2.0]
    m = np.array([[0, 1], [2, 5]], F)
    assert run("sum", [m], {"axes": [0, 1]})[0].ravel().tolist() == [8.0]
    assert run("sum", [m], {"axes": [0]})[0].tolist() == [2.0, 6.0]
    assert run("sum", [m], {"axes": [1]})[0].tolist() == [1.0, 7.0]
    a = np.array([[1, 4], [3, 6]], F)
    assert run("maxElement", [a], {"axes": [0, 1]})[0].ravel().tolist() == [6.0]
    assert run("maxElement", [a], {"axes": [0]})[0].tolist() == [3.0, 6.0]
    assert run("maxElement", [a], {"axes": [1]})[0].tolist() == [4.0, 6.0]
    assert run("argmin", [np.array([4, 2, 6, 1, 2], F)], {"axis": 0})[0].ravel().tolist() == [3]
    assert run("argmin", [np.array([[5, 1, 3], [6, 7, 2]], F)], {"axis": 1})[0].ravel().tolist() == [1, 2]


@pytest.mark.parametrize("shape,axes", [
    ((640, 640, 3, 3), [0, 1, 2, 3]),     # weight-decay term sum(W*W) of the largest WRN-28-10 filter
    ((128, 100), [0, 1]),                 # cross-entropy
    ((6, 7, 8), [1]), ((6, 7, 8), [0, 2]), ((6, 7, 8), [2, 0]), ((5, 3, 4, 2), [1, 3]),
])
def test_sum_and_max(shape, axes):
    x = np.random.RandomState(12).randn(*shape).astype(F)
    got, ref = run("sum", [x], {"axes": axes})
    assert got.shape == ref.shape
    assert rel_err(got, ref) < 1e-4 * max(1.0, np.sqrt(x.size / max(ref.size, 1)) / 50)
    got, ref = run("maxElement", [x], {"axes": axes})
    np.testing.assert_array_equal(got, ref)


def test_sum_int_exact():
    x = np.random.RandomState(13).randint(-100, 100, size=(33, 1000)).astype(np.int32)
    got, ref = run("sum", [x], {"axes": [1]})
    np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_argmin(axis):
    x = np.random.RandomState(14).randn(9, 17, 5).astype(F)
    x[2, 3, 1] = x[4, 3, 1] = -100.0  # tie along axis 0: first wins
    got, ref = run("argmin", [x], {"axis": axis})
    np.testing.assert_array_equal(got, ref)


# ---------------------------------------------------------------------------------------------------------------------
# matmul: the degenerate shapes dopt's lowering produces + real GEMMs
# ---------------------------------------------------------------------------------------------------------------------
def test_matmul_reference_kat():
    got, _ = run("matmul", [np.array([[1], [2]], F), np.array([[3, 4]], F)])
    assert got.ravel().tolist() == [3, 4, 6, 8]


@pytest.mark.parametrize("M,K,N", [
    (3686400, 1, 1),      # scalar broadcast of a [640,640,3,3] tensor: exact
    (128, 1, 100),        # bias broadcast [N,1]x[1,out]: exact
    (1001, 1, 3),
])
def test_matmul_outer_is_exact(M, K, N):
    rng = np.random.RandomState(15)
    a, b = rng.randn(M, K).astype(F), rng.randn(K, N).astype(F)
    got, ref = run("matmul", [a, b])
    np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize("M,K,N,math", [
    (81920, 64, 1, db.MATH_FP32),      # meanPool row sums of WRN-28-10
    (1, 128, 100, db.MATH_FP32),       # column sums (bias gradient of dense)
    (1, 5000, 37, db.MATH_FP32),
    (100, 784, 10, db.MATH_FP32),      # MNIST logistic regression
    (128, 640, 100, db.MATH_FP32),     # WRN dense layer
    (100, 512, 512, db.MATH_FP32),     # VGG dense
    (65, 33, 17, db.MATH_FP32),
])
def test_matmul_fp32(M, K, N, math):
    rng = np.random.RandomState(16)
    a, b = rng.randn(M, K).astype(F), rng.randn(K, N).astype(F)
    got, ref = run("matmul", [a, b], math=math)
    assert rel_err(got, ref) < 1e-5 * max(1.0, np.sqrt(K))


@pytest.mark.parametrize("M,K,N", [(256, 512, 256), (1000, 520, 264), (128, 4096, 64), (300, 200, 650),
                                   (128, 640, 100), (640, 128, 100), (128, 100, 640)])   # + the WRN dense layer and its gradients
def test_matmul_tensor_core(M, K, N):
    rng = np.random.RandomState(17)
    a, b = rng.randn(M, K).astype(F), rng.randn(K, N).astype(F)
    got, ref = run("matmul", [a, b], math=db.MATH_BF16)
    assert rel_err(got, ref) < 2e-2


# ---------------------------------------------------------------------------------------------------------------------
# nnet: relu / bias / pooling / softmax
# ---------------------------------------------------------------------------------------------------------------------
def test_nnet_reference_kats():
    x = np.array([1, 2, 4, 3, 5, 3, 2, 2, 0.1, -4, 3, 2, 0, 0, 2, 2], F).reshape(1, 1, 4, 4)
    got, _ = run("maxpool", [x], {"dims": [2, 2]})
    assert got.ravel().tolist() == [F(5), F(4), F(0.1), F(3)]
    got, _ = run("softmax", [np.array([[1, 2, 3, 1, 2]], F)])
    np.testing.assert_allclose(got.ravel(), [0.0674508, 0.18335, 0.498398, 0.0674508, 0.18335], atol=1e-6)


@pytest.mark.parametrize("shape", [(128, 160, 32, 32), (7, 3, 5, 5), (100, 512)])
def test_relu_and_grad(shape):
    rng = np.random.RandomState(18)
    x = rng.randn(*shape).astype(F)
    x.ravel()[5] = np.nan
    got, ref = run("relu", [x])
    np.testing.assert_array_equal(got, ref)
    dy = rng.randn(*shape).astype(F)
    got, ref2 = run("reluGrad", [dy, ref, x])
    np.testing.assert_array_equal(got, ref2)


def test_add_bias_and_grad():
    rng = np.random.RandomState(19)
    x = rng.randn(100, 64, 32, 32).astype(F)
    b = rng.randn(64).astype(F)
    got, ref = run("addBias", [x, b])
    np.testing.assert_array_equal(got, ref)
    got, ref = run("addBiasGrad", [x])
    assert rel_err(got, ref) < 1e-4
    x2 = rng.randn(3, 5, 7, 2).astype(F)
    got, ref = run("addBiasGrad", [x2])
    assert rel_err(got, ref) < 1e-5


@pytest.mark.parametrize("shape,dims", [((100, 64, 32, 32), [2, 2]), ((3, 5, 9, 7), [2, 3]), ((100, 32, 24, 24), [2, 2])])
def test_maxpool_and_grad(shape, dims):
    rng = np.random.RandomState(20)
    x = rng.randn(*shape).astype(F)
    got, y = run("maxpool", [x], {"dims": dims})
    np.testing.assert_array_equal(got, y)
    dy = rng.randn(*y.shape).astype(F)
    got, ref = run("maxpoolGrad", [dy, y, x], {"dims": dims}, out_shape=shape)
    np.testing.assert_array_equal(got, ref)


def test_maxpool_grad_ties_after_relu():
    # all-zero windows are common after ReLU; only the first tied element receives the gradient, as in cuDNN
    # (measured in test_cudnn_replay_gpu.py::test_maxpool_grad_tie_rule_is_cudnns)
    x = np.zeros((2, 3, 4, 4), F)
    x[0, 0, 0, 1] = 2.0
    y = R.maxpool(x, [2, 2])
    dy = np.arange(1, y.size + 1, dtype=F).reshape(y.shape)
    got, ref = run("maxpoolGrad", [dy, y, x], {"dims": [2, 2]}, out_shape=x.shape)
    np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize("shape", [(128, 100), (100, 10), (4, 7, 3, 5)])
def test_softmax_and_grad(shape):
    rng = np.random.RandomState(21)
    x = (rng.randn(*shape) * 3).astype(F)
    got, y = run("softmax", [x])
    np.testing.assert_allclose(got, y, rtol=2e-6, atol=1e-7)
    dy = rng.randn(*shape).astype(F)
    got, ref = run("softmaxGrad", [dy, y])
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-6)


# ---------------------------------------------------------------------------------------------------------------------
# batch norm
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(128, 160, 32, 32), (128, 640, 8, 8), (48, 64, 16, 16), (5, 3, 7, 5), (100, 512),
                                   (3, 2)])
def test_batchnorm_train_grad_inference(shape):
    rng = np.random.RandomState(22)
    C = shape[1]
    x = (rng.randn(*shape) * 1.7 + 0.8).astype(F)
    g = (rng.rand(C) + 0.5).astype(F)
    b = rng.randn(C).astype(F)
    rm = rng.randn(C).astype(F)
    rv = (rng.rand(C) + 0.5).astype(F)
    gshape = (1, C, 1, 1)
    V = x.size
    got, ref = run("batchNormTrain", [x, g.reshape(gshape), b, rm, rv], {"momentum": 0.9})
    assert got.shape == (V + 2 * C,)
    np.testing.assert_allclose(got[:V], ref[:V], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(got[V:], ref[V:], rtol=1e-5, atol=1e-6)
    dy = rng.randn(*shape).astype(F)
    # the reference over-allocates the result to vol(dy)+vol(x)+vol(scale) (core/ops/nnet.d:232-235); the tail stays zero
    got, ref = run("batchNormGrad", [dy, x, g.reshape(gshape)], out_shape=(2 * V + C,))
    np.testing.assert_allclose(got[:V], ref[:V], rtol=1e-3, atol=1e-4 * float(np.abs(ref[:V]).max() + 1e-6))
    np.testing.assert_allclose(got[V:V + 2 * C], ref[V:V + 2 * C], rtol=2e-4,
                               atol=2e-4 * float(np.abs(ref[V:V + 2 * C]).max()))
    assert not got[V + 2 * C:].any()
    got, ref = run("batchNormInference", [x, g.reshape(gshape), b, rm, rv])
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-5)


def test_batchnorm_large_mean_is_stable():
    rng = np.random.RandomState(23)
    x = (rng.randn(64, 8, 16, 16) * 0.01 + 100.0).astype(F)
    one, zero = np.ones(8, F), np.zeros(8, F)
    got, ref = run("batchNormTrain", [x, one.reshape(1, 8, 1, 1), zero, zero, one], {"momentum": 0.9})
    np.testing.assert_allclose(got[:x.size], ref[:x.size], rtol=0, atol=2e-2)
    np.testing.assert_allclose(got[x.size:], ref[x.size:], rtol=1e-3, atol=1e-5)


def test_batchnorm_running_mean_kat():
    # nnet/source/dopt/nnet/layers/batchnorm.d:158-177
    x = np.array([[1, 2], [3, 4], [5, 6]], F)
    one, zero = np.ones(2, F), np.zeros(2, F)
    rm, rv = zero.copy(), one.copy()
    for _ in range(200):
        got, _ = run("batchNormTrain", [x, one, zero, rm, rv], {"momentum": 0.9})
        rm, rv = got[6:8].copy(), got[8:10].copy()
    np.testing.assert_allclose(rm, [3.0, 4.0], rtol=1e-2, atol=1e-5)


# ---------------------------------------------------------------------------------------------------------------------
# convolution family
# ---------------------------------------------------------------------------------------------------------------------
def test_convolution_reference_kat_both_paths():
    x = np.array([1, 1, 1, 0, 0] * 3, F).reshape(1, 1, 3, 5)
    w = np.array([-1, 1], F).reshape(1, 1, 1, 2)
    for math in (db.MATH_FP32, db.MATH_BF16):
        got, _ = run("convolution", [x, w], {"padding": [0, 0], "stride": [1, 1]}, math=math)
        assert got.ravel().tolist() == [0, 0, 1, 0] * 3


CONV_CASES = [
    # N, C, H, W, K, R, S, pad, stride
    (2, 3, 8, 8, 4, 3, 3, 1, 1),
    (4, 16, 32, 32, 16, 3, 3, 1, 1),       # WRN stem-like
    (8, 32, 16, 16, 64, 3, 3, 1, 1),
    (8, 64, 16, 16, 32, 3, 3, 1, 2),       # stride 2
    (8, 64, 16, 16, 128, 1, 1, 0, 2),      # 1x1 stride-2 shortcut
    (8, 16, 16, 16, 160, 1, 1, 0, 1),      # 1x1 shortcut
    (16, 160, 8, 8, 160, 3, 3, 1, 1),      # C not a multiple of 64
    (4, 320, 8, 8, 320, 3, 3, 1, 1),       # two N tiles of 160
    (100, 1, 28, 28, 32, 5, 5, 0, 1),      # MNIST conv1 (unpadded 5x5)
    (10, 32, 12, 12, 32, 5, 5, 0, 1),      # MNIST conv2
    (2, 64, 24, 24, 64, 3, 3, 1, 1),       # SINS-like non-power-of-two width
    (3, 64, 12, 12, 64, 3, 3, 1, 2),
    (32, 128, 4, 4, 128, 3, 3, 1, 1),      # VGG 4x4
    (64, 256, 2, 2, 256, 3, 3, 1, 1),      # VGG 2x2
]


@pytest.mark.parametrize("math,tol", [(db.MATH_FP32, 1e-4), (db.MATH_BF16, 2e-2)])
@pytest.mark.parametrize("case", CONV_CASES)
def test_convolution_family(case, math, tol):
    N, C, H, W, K, Rr, S, pad, stride = case
    rng = np.random.RandomState(24)
    x = rng.randn(N, C, H, W).astype(F)
    w = (rng.randn(K, C, Rr, S) * np.sqrt(2.0 / (C * Rr * S))).astype(F)
    attrs = {"padding": [pad, pad], "stride": [stride, stride]}
    y_got, y = run("convolution", [x, w], attrs, math=math)
    assert y_got.shape == y.shape
    assert rel_err(y_got, y) < tol, "fwd"
    dy = rng.randn(*y.shape).astype(F)
    a = dict(attrs, featuresShape=list(x.shape))
    dx_got, dx = run("convolutionFeaturesGrad", [dy, w], a, math=math, out_shape=x.shape)
    assert rel_err(dx_got, dx) < tol, "dgrad"
    a = dict(attrs, filtersShape=list(w.shape))
    dw_got, dw = run("convolutionFiltersGrad", [dy, x], a, math=math, out_shape=w.shape)
    assert rel_err(dw_got, dw) < tol, "wgrad"


@pytest.mark.parametrize("case", [
    (128, 160, 32, 32, 160, 3, 3, 1, 1),   # the three families that carry 94 % of WRN-28-10's FLOPs (full size)
    (128, 320, 16, 16, 320, 3, 3, 1, 1),
    (128, 640, 8, 8, 640, 3, 3, 1, 1),
    (128, 160, 32, 32, 320, 3, 3, 1, 2),
])
def test_convolution_full_size_vs_torch(case):
    """At BASELINE sizes the numpy oracle is too slow for the test budget, so the check is against torch's fp32 conv
    (cuDNN, TF32 off) -- the same library family the reference calls -- plus a linearity property."""
    N, C, H, W, K, Rr, S, pad, stride = case
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(25)
    x = torch.randn(N, C, H, W, device="cuda", generator=g)
    w = torch.randn(K, C, Rr, S, device="cuda", generator=g) * (2.0 / (C * Rr * S)) ** 0.5
    P = (H + 2 * pad - Rr) // stride + 1
    y = db.run_op("convolution", [x, w], (N, K, P, P), padding=[pad, pad], stride=[stride, stride])
    ref = torch.nn.functional.conv2d(x, w.flip(2, 3), padding=pad, stride=stride)   # true convolution = flipped xcorr
    assert float((y - ref).abs().max() / ref.abs().max()) < 2e-2
    dy = torch.randn_like(ref)
    dx = db.run_op("convolutionFeaturesGrad", [dy, w], (N, C, H, W), padding=[pad, pad], stride=[stride, stride])
    dw = db.run_op("convolutionFiltersGrad", [dy, x], (K, C, Rr, S), padding=[pad, pad], stride=[stride, stride])
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    torch.nn.functional.conv2d(xr, wr.flip(2, 3), padding=pad, stride=stride).backward(dy)
    assert float((dx - xr.grad).abs().max() / xr.grad.abs().max()) < 2e-2
    assert float((dw - wr.grad).abs().max() / wr.grad.abs().max()) < 2e-2
    # linearity in the features: conv(2x) == 2 conv(x) exactly (scaling by 2 commutes with bf16 rounding)
    y2 = db.run_op("convolution", [2 * x, w], (N, K, P, P), padding=[pad, pad], stride=[stride, stride])
    assert torch.equal(y2, 2 * y)


# ---------------------------------------------------------------------------------------------------------------------
# uniform
# ---------------------------------------------------------------------------------------------------------------------
def test_uniform_distribution():
    out = db.run_op("uniform", [], (1 << 20,), seed=1234)
    v = host(out)
    assert v.min() > 0.0 and v.max() <= 1.0            # (0, 1] like curandGenerateUniform
    assert abs(v.mean() - 0.5) < 2e-3 and abs(v.var() - 1.0 / 12) < 2e-3
    out2 = db.run_op("uniform", [], (1 << 20,), seed=1234)
    assert np.array_equal(v, host(out2))               # seeded: reproducible


@pytest.mark.parametrize("math", [db.MATH_FP32, db.MATH_BF16], ids=["fp32", "bf16"])
def test_kernel_handle_sees_weights_overwritten_in_place(math):
    """dopt.online overwrites a parameter's CUDABuffer in place after every step (online/source/dopt/online/sgd.d:76-91 ->
    cuda/source/dopt/cuda/package.d:419-422) and then executes the SAME CUDAKernel objects again.  A kernel handle must
    therefore never reuse anything derived from an earlier value of an operand (packed / staged filters): the second execute
    on the same handle, with the filter buffer overwritten at the same address, has to give the new result -- for the
    forward convolution, its feature gradient (filter operand) and its filter gradient (activation operands)."""
    rng = np.random.RandomState(11)
    N, C, H, W, K = 8, 32, 16, 16, 64
    x = rng.randn(N, C, H, W).astype(F)
    dy = rng.randn(N, K, H, W).astype(F)
    w1 = (rng.randn(K, C, 3, 3) * 0.1).astype(F)
    w2 = (rng.randn(K, C, 3, 3) * 0.1).astype(F)
    attrs = dict(padding=[1, 1], stride=[1, 1])
    tol = 1e-4 if math == db.MATH_FP32 else 2e-2
    s = torch.cuda.current_stream().cuda_stream

    def close(got, ref):
        return float(np.abs(got - ref).max()) <= tol * float(np.abs(ref).max())

    xd, dyd, wd = dev(x), dev(dy), dev(w1)
    fwd = db.CUDAKernel(db.make_op("convolution", [x.shape, w1.shape], (N, K, H, W), None, db.FLOAT32, math, **attrs))
    dgr = db.CUDAKernel(db.make_op("convolutionFeaturesGrad", [dy.shape, w1.shape], x.shape, None, db.FLOAT32, math,
                                   featuresShape=list(x.shape), **attrs))
    wgr = db.CUDAKernel(db.make_op("convolutionFiltersGrad", [dy.shape, x.shape], w1.shape, None, db.FLOAT32, math,
                                   filtersShape=list(w1.shape), **attrs))
    y = torch.zeros((N, K, H, W), device="cuda")
    dx = torch.zeros(x.shape, device="cuda")
    dw = torch.zeros(w1.shape, device="cuda")
    for w_now, x_now in ((w1, x), (w2, x * 0.5 + 1.0)):
        wd.copy_(dev(w_now))            # same device address, new value
        xd.copy_(dev(x_now.astype(F)))
        fwd.execute([xd, wd], y, s)
        dgr.execute([dyd, wd], dx, s)
        wgr.execute([dyd, xd], dw, s)
        assert close(host(y), R.convolution(x_now.astype(F), w_now, (1, 1), (1, 1)))
        assert close(host(dx), R.convolution_features_grad(dy, w_now, x.shape, (1, 1), (1, 1)))
        assert close(host(dw), R.convolution_filters_grad(dy, x_now.astype(F), w1.shape, (1, 1), (1, 1)))
    for k in (fwd, dgr, wgr):
        k.close()
