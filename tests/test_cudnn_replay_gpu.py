"""Parity against the reference CUDA backend's own arithmetic: the cuDNN / cuBLAS call sequences of
cuda/source/dopt/cuda/nnet/cudnn7.d, math.d and basic.d, replayed from C++ (oracle/cudnn_replay.cpp) on the same device
and the same inputs as the product kernels, which are called through the C ABI like CUDAKernel.execute.

Three-way check per op: product == cuDNN replay (the north star's "conv/BN/pool against the reference CUDA backend"), and
cuDNN replay == numpy oracle at small sizes (which pins the oracle's restatement of the ops the reference CPU backend
lacks: conv gradients, maxpoolGrad, softmaxGrad, relu(+Grad), addBias(+Grad), batchNorm*).

Tolerances:
  * relu / reluGrad / maxpool / maxpoolGrad / addBias / transpose: bit-exact
  * softmax(+Grad), bias gradient, fp32 matmul: 1e-4 of the tensor's max magnitude (summation order); batch norm: 2e-4
  * convolution family, product MATH_FP32 vs cuDNN with CUDNN_FMA_MATH: 1e-4
  * convolution family, product MATH_BF16 (tcgen05) vs cuDNN as the reference configures it (default math type -- cuDNN
    may use TF32 tensor cores for fp32 data on this GPU): 2e-2

With DOPT_B200_REPORT=<file> every measured error is appended to that file as a JSON line (profiles/ keeps a copy).
"""
import json
import os

import numpy as np
import pytest
import torch

import dopt_b200 as db
from oracle import cudnn_replay as Q
from oracle import dopt_ref as R

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not Q.available(), reason="oracle/_build/libcudnn_replay.so not built")]
F = np.float32


def report(name, **kv):
    path = os.environ.get("DOPT_B200_REPORT")
    if path:
        with open(path, "a") as f:
            f.write(json.dumps(dict(test=name, **kv)) + "\n")


def rel(got, ref):
    scale = max(float(ref.abs().max()), 1e-30)
    return float((got.double() - ref.double()).abs().max()) / scale


def rand(shape, seed, scale=1.0, shift=0.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(*shape, device="cuda", generator=g) * scale + shift


def urand(n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.rand(n, device="cuda", generator=g) + 0.5


def test_versions():
    v = Q.versions()
    report("versions", **v)
    assert v["cudnn"] >= 9000


# ---------------------------------------------------------------------------------------------------------------------
# convolution family
# ---------------------------------------------------------------------------------------------------------------------
def test_replay_reproduces_the_reference_convolution_kat():
    # core/source/dopt/core/ops/nnet.d:270-293 -- cuDNN in CUDNN_CONVOLUTION mode flips the filter
    x = torch.tensor([1, 1, 1, 0, 0] * 3, dtype=torch.float32, device="cuda").reshape(1, 1, 3, 5)
    w = torch.tensor([-1, 1], dtype=torch.float32, device="cuda").reshape(1, 1, 1, 2)
    y = Q.convolution(x, w, [0, 0], [1, 1], math=1)
    assert y.flatten().tolist() == [0, 0, 1, 0] * 3


SMALL_CONV = [
    # N, C, H, W, K, R, S, pad, stride
    (2, 3, 8, 8, 4, 3, 3, 1, 1),
    (3, 8, 9, 7, 6, 3, 2, 1, 2),
    (10, 1, 28, 28, 32, 5, 5, 0, 1),
    (4, 16, 12, 12, 16, 1, 1, 0, 2),
]


@pytest.mark.parametrize("case", SMALL_CONV)
def test_numpy_oracle_matches_cudnn_for_the_convolution_family(case):
    N, C, H, W, K, Rr, S, pad, stride = case
    x = rand((N, C, H, W), 1)
    w = rand((K, C, Rr, S), 2, (2.0 / (C * Rr * S)) ** 0.5)
    p, s = [pad, pad], [stride, stride]
    y = Q.convolution(x, w, p, s, math=1)
    yr = torch.from_numpy(R.convolution(x.cpu().numpy(), w.cpu().numpy(), p, s)).cuda()
    assert y.shape == yr.shape
    dy = rand(tuple(y.shape), 3)
    dx = Q.convolution_features_grad(dy, w, x.shape, p, s, math=1)
    dw = Q.convolution_filters_grad(dy, x, w.shape, p, s, math=1)
    dxr = torch.from_numpy(R.convolution_features_grad(dy.cpu().numpy(), w.cpu().numpy(), tuple(x.shape), p, s)).cuda()
    dwr = torch.from_numpy(R.convolution_filters_grad(dy.cpu().numpy(), x.cpu().numpy(), tuple(w.shape), p, s)).cuda()
    errs = dict(fwd=rel(yr, y), dgrad=rel(dxr, dx), wgrad=rel(dwr, dw))
    report("oracle_vs_cudnn_conv", case=list(case), **errs)
    assert max(errs.values()) < 1e-4, errs


CONV = [
    (8, 32, 16, 16, 64, 3, 3, 1, 1),
    (8, 64, 16, 16, 32, 3, 3, 1, 2),
    (8, 64, 16, 16, 128, 1, 1, 0, 2),
    (16, 160, 8, 8, 160, 3, 3, 1, 1),
    (100, 1, 28, 28, 32, 5, 5, 0, 1),
    (128, 3, 32, 32, 16, 3, 3, 1, 1),      # WRN stem at BASELINE size
    (128, 160, 32, 32, 160, 3, 3, 1, 1),   # the three WRN-28-10 families at BASELINE size
    (128, 320, 16, 16, 320, 3, 3, 1, 1),
    (128, 640, 8, 8, 640, 3, 3, 1, 1),
    (128, 160, 32, 32, 320, 3, 3, 1, 2),
    (128, 160, 32, 32, 320, 1, 1, 0, 2),
    (128, 16, 32, 32, 160, 3, 3, 1, 1),    # the remaining WRN-28-10 shapes of SURVEY section 8(d) at BASELINE size
    (128, 16, 32, 32, 160, 1, 1, 0, 1),
    (128, 320, 16, 16, 640, 3, 3, 1, 2),
    (128, 320, 16, 16, 640, 1, 1, 0, 2),
    # BASELINE.json configs[2], VGG19 on 100 x 3 x 32 x 32 (nnet/models/vgg.d; examples/cifar10.d:54-66): every distinct shape.
    # The 4 x 4 and 2 x 2 feature maps pack 8 / 32 images into one 128-pixel tile.
    (100, 3, 32, 32, 64, 3, 3, 1, 1),
    (100, 64, 32, 32, 64, 3, 3, 1, 1),
    (100, 64, 16, 16, 128, 3, 3, 1, 1),
    (100, 128, 16, 16, 128, 3, 3, 1, 1),
    (100, 128, 8, 8, 256, 3, 3, 1, 1),
    (100, 256, 8, 8, 256, 3, 3, 1, 1),
    (100, 256, 4, 4, 512, 3, 3, 1, 1),
    (100, 512, 4, 4, 512, 3, 3, 1, 1),
    (100, 512, 2, 2, 512, 3, 3, 1, 1),
    # configs[4], Wide ResNet-16-8 on 50 x 3 x 96 x 96 with strides [2,2,2] (examples/sins10.d:42-52): 96- and 48-pixel rows
    # (tiles that 128 pixels do not divide), 24 x 24 and 12 x 12 maps, three stride-2 transitions with their 1 x 1 shortcuts
    (50, 3, 96, 96, 16, 3, 3, 1, 1),
    (50, 16, 96, 96, 128, 3, 3, 1, 2),
    (50, 16, 96, 96, 128, 1, 1, 0, 2),
    (50, 128, 48, 48, 128, 3, 3, 1, 1),
    (50, 128, 48, 48, 256, 3, 3, 1, 2),
    (50, 128, 48, 48, 256, 1, 1, 0, 2),
    (50, 256, 24, 24, 256, 3, 3, 1, 1),
    (50, 256, 24, 24, 512, 3, 3, 1, 2),
    (50, 256, 24, 24, 512, 1, 1, 0, 2),
    (50, 512, 12, 12, 512, 3, 3, 1, 1),
    # configs[1], the MNIST CNN's second 5 x 5 unpadded convolution (examples/mnist.d; SURVEY 8a: M=6400, N=32, K=800)
    (100, 32, 12, 12, 32, 5, 5, 0, 1),
]


@pytest.mark.parametrize("math,tol", [(db.MATH_FP32, 1e-4), (db.MATH_BF16, 2e-2)], ids=["fp32", "bf16"])
@pytest.mark.parametrize("case", CONV)
def test_convolution_family_vs_cudnn(case, math, tol):
    N, C, H, W, K, Rr, S, pad, stride = case
    if math == db.MATH_FP32 and N * C * H * W > (1 << 22):
        pytest.skip("fp32 SIMT path at BASELINE size: covered at the smaller shapes")
    x = rand((N, C, H, W), 4)
    w = rand((K, C, Rr, S), 5, (2.0 / (C * Rr * S)) ** 0.5)
    p, s = [pad, pad], [stride, stride]
    qmath = 1 if math == db.MATH_FP32 else 0
    y = Q.convolution(x, w, p, s, math=qmath)
    dy = rand(tuple(y.shape), 6)
    dx = Q.convolution_features_grad(dy, w, x.shape, p, s, math=qmath)
    dw = Q.convolution_filters_grad(dy, x, w.shape, p, s, math=qmath)
    y_g = db.run_op("convolution", [x, w], tuple(y.shape), math=math, padding=p, stride=s)
    dx_g = db.run_op("convolutionFeaturesGrad", [dy, w], tuple(x.shape), math=math, padding=p, stride=s)
    dw_g = db.run_op("convolutionFiltersGrad", [dy, x], tuple(w.shape), math=math, padding=p, stride=s)
    errs = dict(fwd=rel(y_g, y), dgrad=rel(dx_g, dx), wgrad=rel(dw_g, dw))
    report("conv_vs_cudnn", case=list(case), math="fp32" if math == db.MATH_FP32 else "bf16", **errs)
    assert max(errs.values()) < tol, errs


def test_cudnn_default_math_error_on_this_gpu():
    """What the unmodified reference would compute on this GPU: cuDNN's default math type against strict fp32 FMA.  Not a
    product check -- it records the precision of the parity target itself (reported next to the bf16 path's error)."""
    x = rand((128, 160, 32, 32), 7)
    w = rand((160, 160, 3, 3), 8, (2.0 / 1440) ** 0.5)
    strict = Q.convolution(x, w, [1, 1], [1, 1], math=1)
    default = Q.convolution(x, w, [1, 1], [1, 1], math=0)
    ours = db.run_op("convolution", [x, w], tuple(strict.shape), math=db.MATH_BF16, padding=[1, 1], stride=[1, 1])
    e_def, e_ours = rel(default, strict), rel(ours, strict)
    report("reference_default_math", cudnn_default_vs_fma=e_def, dopt_b200_bf16_vs_fma=e_ours)
    assert e_ours < 2e-2


# ---------------------------------------------------------------------------------------------------------------------
# pooling: values and tie routing
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape,dims", [((100, 64, 32, 32), [2, 2]), ((3, 5, 9, 7), [2, 3]), ((100, 32, 24, 24), [2, 2]),
                                        ((100, 512, 4, 4), [2, 2]), ((100, 512, 2, 2), [2, 2])])   # VGG19's last two pools
def test_maxpool_and_grad_vs_cudnn(shape, dims):
    x = rand(shape, 9)
    y = Q.maxpool(x, dims)
    y_g = db.run_op("maxpool", [x], tuple(y.shape), dims=dims)
    assert torch.equal(y_g, y)
    assert np.array_equal(y.cpu().numpy(), R.maxpool(x.cpu().numpy(), dims))
    dy = rand(tuple(y.shape), 10)
    dx = Q.maxpool_grad(dy, y, x, dims)
    dx_g = db.run_op("maxpoolGrad", [dy, y, x], shape, dims=dims)
    assert torch.equal(dx_g, dx)


def test_maxpool_grad_tie_rule_is_cudnns(monkeypatch):
    """All-zero windows are the common case after ReLU.  Measures which elements of a tied window cuDNN's PoolingBackward
    routes dy to and requires the product's default rule to give the same result."""
    x = torch.zeros((2, 3, 4, 4), device="cuda")
    x[0, 0, 0, 1] = 2.0
    x[1, 2, 2:4, 2:4] = 1.5                                   # a fully tied non-zero window
    x[1, 1, 0, 0] = -1.0                                      # window max 0 shared by three elements
    y = Q.maxpool(x, [2, 2])
    dy = torch.arange(1, y.numel() + 1, dtype=torch.float32, device="cuda").reshape(y.shape)
    ref = Q.maxpool_grad(dy, y, x, [2, 2])
    tied_window = ref[1, 2, 2:4, 2:4].flatten().tolist()
    rule = "all" if all(v == tied_window[0] and v != 0 for v in tied_window) else \
        "first" if tied_window[0] != 0 and not any(tied_window[1:]) else "other"
    got = {}
    for mode in ("all", "first"):
        monkeypatch.setenv("DOPT_B200_POOL_TIES", mode)
        got[mode] = db.run_op("maxpoolGrad", [dy, y, x], tuple(x.shape), dims=[2, 2])
    monkeypatch.delenv("DOPT_B200_POOL_TIES")
    default = db.run_op("maxpoolGrad", [dy, y, x], tuple(x.shape), dims=[2, 2])
    report("maxpool_grad_ties", cudnn_rule=rule, tied_window=tied_window,
           matches_all=bool(torch.equal(got["all"], ref)), matches_first=bool(torch.equal(got["first"], ref)),
           default_matches=bool(torch.equal(default, ref)))
    assert torch.equal(default, ref), "cuDNN routes ties by rule '%s'" % rule
    np.testing.assert_array_equal(R.maxpool_grad(dy.cpu().numpy(), y.cpu().numpy(), x.cpu().numpy(), [2, 2]),
                                  ref.cpu().numpy())


# ---------------------------------------------------------------------------------------------------------------------
# softmax / relu / bias
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(128, 100), (100, 10), (4, 7, 3, 5)])
def test_softmax_and_grad_vs_cudnn(shape):
    x = rand(shape, 11, 3.0)
    y = Q.softmax(x)
    y_g = db.run_op("softmax", [x], shape)
    dy = rand(shape, 12)
    dx = Q.softmax_grad(dy, y)
    dx_g = db.run_op("softmaxGrad", [dy, y], shape)
    yo = torch.from_numpy(R.softmax(x.cpu().numpy())).cuda()
    dxo = torch.from_numpy(R.softmax_grad(dy.cpu().numpy(), y.cpu().numpy())).cuda()
    errs = dict(fwd=rel(y_g, y), grad=rel(dx_g, dx), oracle_fwd=rel(yo, y), oracle_grad=rel(dxo, dx))
    report("softmax_vs_cudnn", shape=list(shape), **errs)
    assert max(errs.values()) < 1e-4, errs


@pytest.mark.parametrize("shape", [(128, 160, 32, 32), (7, 3, 5, 5), (100, 512)])
def test_relu_and_grad_vs_cudnn(shape):
    x = rand(shape, 13)
    x.view(-1)[::7] = 0.0                                     # exact zeros: gate is x > 0
    y = Q.relu(x)
    y_g = db.run_op("relu", [x], shape)
    assert torch.equal(y_g, y)
    dy = rand(shape, 14)
    dx = Q.relu_grad(dy, y, x)
    dx_g = db.run_op("reluGrad", [dy, y, x], shape)
    assert torch.equal(dx_g, dx)
    assert np.array_equal(R.relu_grad(dy.cpu().numpy(), y.cpu().numpy(), x.cpu().numpy()), dx.cpu().numpy())


def test_add_bias_and_grad_vs_cudnn():
    x = rand((100, 64, 16, 16), 15)
    b = rand((64,), 16)
    out = Q.add_bias(x, b)
    out_g = db.run_op("addBias", [x, b], tuple(x.shape))
    assert torch.equal(out_g, out)
    dy = rand((100, 64, 16, 16), 17)
    db1 = Q.add_bias_grad(dy)
    db_g = db.run_op("addBiasGrad", [dy], (64,))
    err = rel(db_g, db1)
    # the reference accumulates across executions (beta = 1 into a buffer zeroed once); the product
    # implements the first-execution value every time
    db2 = Q.add_bias_grad(dy, executions=2)
    report("add_bias_grad", err=err, reference_second_execution_ratio=float((db2 / db1).mean()))
    assert err < 1e-4
    assert rel(db2, 2 * db1) < 1e-5


# ---------------------------------------------------------------------------------------------------------------------
# batch norm
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(128, 160, 32, 32), (128, 640, 8, 8), (48, 64, 16, 16), (5, 3, 7, 5), (100, 512),
                                   (100, 512, 2, 2), (100, 64, 32, 32), (50, 128, 48, 48), (50, 512, 12, 12)])   # VGG19 / SINS
def test_batchnorm_vs_cudnn(shape):
    C = shape[1]
    x = rand(shape, 18, 1.7, 0.8)
    g = urand(C, 30)
    b = rand((C,), 19)
    rm = rand((C,), 20)
    rv = urand(C, 31)
    V = x.numel()
    gs = g.reshape(1, C, 1, 1)                                # the layer's gamma is [1, C, 1, 1] (layers/batchnorm.d:75)
    ref = Q.batch_norm_train(x, g, b, rm, rv, 0.9)
    got = db.run_op("batchNormTrain", [x, gs, b, rm, rv], (V + 2 * C,), momentum=0.9)
    e_y, e_stats = rel(got[:V], ref[:V]), rel(got[V:], ref[V:])
    dy = rand(shape, 21)
    refg = Q.batch_norm_grad(dy, x, g)
    gotg = db.run_op("batchNormGrad", [dy, x, gs], (2 * V + C,))
    e_dx, e_dp = rel(gotg[:V], refg[:V]), rel(gotg[V:V + 2 * C], refg[V:V + 2 * C])
    refi = Q.batch_norm_inference(x, g, b, rm, rv)
    goti = db.run_op("batchNormInference", [x, gs, b, rm, rv], shape)
    e_inf = rel(goti, refi)
    errs = dict(y=e_y, running=e_stats, dx=e_dx, dscale_dbias=e_dp, inference=e_inf)
    report("batchnorm_vs_cudnn", shape=list(shape), **errs)
    assert max(errs.values()) < 2e-4, errs
    assert not gotg[V + 2 * C:].any()


def test_numpy_oracle_matches_cudnn_for_batchnorm():
    shape, C = (6, 5, 7, 3), 5
    x = rand(shape, 22, 1.3, -0.4)
    g = urand(C, 32)
    b, rm = rand((C,), 23), rand((C,), 24)
    rv = urand(C, 33)
    n = lambda t: t.cpu().numpy()
    ref = Q.batch_norm_train(x, g, b, rm, rv, 0.9)
    ora = torch.from_numpy(R.batch_norm_train(n(x), n(g), n(b), n(rm), n(rv), 0.9)).cuda()
    assert rel(ora, ref) < 1e-5
    dy = rand(shape, 25)
    refg = Q.batch_norm_grad(dy, x, g)
    orag = torch.from_numpy(R.batch_norm_grad(n(dy), n(x), n(g), 2 * x.numel() + C)).cuda()
    assert rel(orag[:x.numel()], refg[:x.numel()]) < 1e-4
    assert rel(orag[x.numel():x.numel() + 2 * C], refg[x.numel():x.numel() + 2 * C]) < 1e-4
    refi = Q.batch_norm_inference(x, g, b, rm, rv)
    orai = torch.from_numpy(R.batch_norm_inference(n(x), n(g), n(b), n(rm), n(rv))).cuda()
    assert rel(orai, refi) < 1e-5


# ---------------------------------------------------------------------------------------------------------------------
# cuBLAS: matmul (math.d:214-247) and transpose (basic.d:219-247)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,K,N", [(128, 640, 100), (100, 784, 10), (100, 512, 512), (81920, 64, 1), (4096, 1, 1),
                                   (128, 1, 100)])
def test_matmul_fp32_vs_cublas(M, K, N):
    a, b = rand((M, K), 26), rand((K, N), 27)
    ref = Q.matmul(a, b)
    got = db.run_op("matmul", [a, b], (M, N), math=db.MATH_FP32)
    err = rel(got, ref)
    report("matmul_vs_cublas", shape=[M, K, N], err=err)
    assert err < 1e-4


@pytest.mark.parametrize("rows,cols", [(100, 640), (640, 100), (33, 65)])
def test_transpose_vs_cublas(rows, cols):
    a = rand((rows, cols), 28)
    ref = Q.transpose(a)
    got = db.run_op("transpose", [a], (cols, rows), order=[1, 0])
    assert torch.equal(got, ref)
    assert torch.equal(ref, a.t().contiguous())
