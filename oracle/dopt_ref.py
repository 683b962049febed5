"""CPU oracle for dopt's operations -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.  The product
(libdopt_b200.so / libdopt_host.so) never calls into it and has no CPU fallback.

What it restates (all citations relative to the dopt tree, /root/reference in the build container):
  * the reference CPU backend, op for op: pointwise (cpu/source/dopt/cpu/math.d:314-437), matmul (math.d:70-88, cblas
    sgemm -> numpy/OpenBLAS sgemm), sum / maxElement / argmin (math.d:90-310), slice / pad / transpose / repeat
    (cpu/source/dopt/cpu/basic.d), convolution / maxpool / softmax (cpu/source/dopt/cpu/nnet.d);
  * for ops the CPU backend lacks (conv gradients, maxpoolGrad, softmaxGrad, relu(+Grad), addBias(+Grad), batchNorm*)
    the cuDNN definitions the reference CUDA backend selects (cuda/source/dopt/cuda/nnet/cudnn7.d: CUDNN_CONVOLUTION,
    CUDNN_POOLING_MAX, CUDNN_SOFTMAX_ACCURATE / MODE_CHANNEL, CUDNN_ACTIVATION_RELU, CUDNN_BATCHNORM_SPATIAL, eps 1e-5).
    cuDNN itself is a third-party dependency that is not in the tree (derelict-cudnn ~>2.0.4, cuda/dub.json:3-8; no
    version pins exist because dub.selections.json is git-ignored), so these follow its published formulas.

Deliberate deviation: the reference CPU `convolution` bounds-checks input coordinates against the OUTPUT size
(cpu/source/dopt/cpu/nnet.d:54), which is wrong for unpadded or strided convolutions; the oracle checks against the
input size, i.e. it follows the CUDA backend (cuDNN), which is the path being replaced.

Pinning: tests/test_oracle_golden.py checks this module against every known-answer vector in the reference's own unit
tests (convolution, maxpool, softmax, matmul, sum, argmin, maxElement, slice, pad, reshape, transpose, repeat, d(x*x)/dx,
sliceGrad, the CUDA smoke test and the batch-norm running-mean test).  Ops with no reference test (conv gradients,
maxpoolGrad tie routing, softmaxGrad, relu*, addBias*, batchNormGrad/Inference, optimiser numerics) are "parity
unpinned" by the reference's own tests; for those the oracle is cross-checked against float64 finite differences / closed
forms in the same test file and, on the GPU box, against the reference CUDA backend's own cuDNN / cuBLAS calls replayed
from C++ (oracle/cudnn_replay.cpp, tests/test_cudnn_replay_gpu.py: conv family, pooling incl. tie routing, softmax, relu,
bias, batch norm -- agreement 1e-6 or better, measured errors in profiles/r01p_cudnn_replay_report.jsonl).  Outputs of
those replayed calls are committed as tests/golden/cudnn_golden.npz and checked on CPU by
tests/test_oracle_vs_cudnn_golden.py.

All arithmetic is float32 unless a comment says otherwise (reductions whose order cuDNN does not document accumulate
in float64 and round once; the parity tolerance covers the difference).
"""
import numpy as np

F32 = np.float32
BN_EPS = 1e-5  # cudnn7.d:610,633,653


# --------------------------------------------------------------------------------------------------------------------
# pointwise -- cpu/source/dopt/cpu/math.d:314-437 (semantics), cuda/source/dopt/cuda/math.d:79-170 (CUDA C expressions)
# --------------------------------------------------------------------------------------------------------------------
def _sgn(a):
    return ((a > 0).astype(a.dtype) - (a < 0).astype(a.dtype)).astype(a.dtype)


def _int_via_float(fn):
    def f(a):
        if a.dtype == np.int32:
            with np.errstate(all="ignore"):
                return fn(a.astype(np.float64)).astype(np.int32)  # expCast / sqrtCast, math.d:28-50
        with np.errstate(all="ignore"):
            return fn(a).astype(a.dtype)
    return f


def _div(a, b):
    if a.dtype == np.int32:
        # D / C integer division truncates toward zero; x/0 is undefined in both (library returns 0)
        q = np.zeros_like(a)
        nz = b != 0
        q[nz] = (np.abs(a[nz]) // np.abs(b[nz])) * np.sign(a[nz]) * np.sign(b[nz])
        return q.astype(np.int32)
    with np.errstate(all="ignore"):
        return (a / b).astype(a.dtype)


def _pow(a, b):
    with np.errstate(all="ignore"):
        if a.dtype == np.int32:
            return np.power(a.astype(np.float64), b.astype(np.float64)).astype(np.int32)
        return np.power(a, b).astype(a.dtype)


POINTWISE_BINARY = {
    "add": lambda a, b: (a + b).astype(a.dtype),
    "sub": lambda a, b: (a - b).astype(a.dtype),
    "mul": lambda a, b: (a * b).astype(a.dtype),
    "div": _div,
    "lt": lambda a, b: (a < b).astype(a.dtype),
    "lte": lambda a, b: (a <= b).astype(a.dtype),
    "gt": lambda a, b: (a > b).astype(a.dtype),
    "gte": lambda a, b: (a >= b).astype(a.dtype),
    "eq": lambda a, b: (a == b).astype(a.dtype),
    "neq": lambda a, b: (a != b).astype(a.dtype),
    "max": lambda a, b: np.fmax(a, b).astype(a.dtype),
    "min": lambda a, b: np.fmin(a, b).astype(a.dtype),
    "pow": _pow,
}
POINTWISE_UNARY = {
    "neg": lambda a: (-a).astype(a.dtype),
    "abs": lambda a: np.abs(a).astype(a.dtype),
    "sgn": _sgn,
    "exp": _int_via_float(np.exp),
    "log": _int_via_float(np.log),
    "sqrt": _int_via_float(np.sqrt),
}


def _via_float32(fn):
    """`cast(T)(op(cast(float)x))`, cpu/source/dopt/cpu/math.d:416-423: std.math evaluates the float argument in extended
    precision, so the result is the correctly rounded function of the float32 value (float64 here), then cast to T --
    truncation toward zero for int32."""
    def f(a):
        with np.errstate(all="ignore"):
            r = fn(a.astype(np.float32).astype(np.float64))
            if a.dtype == np.int32:
                return np.trunc(np.nan_to_num(r.astype(np.float32))).astype(np.int32)
            return r.astype(a.dtype)
    return f


for _name, _fn in (("sin", np.sin), ("cos", np.cos), ("tan", np.tan), ("asin", np.arcsin), ("acos", np.arccos),
                   ("atan", np.arctan), ("sinh", np.sinh), ("cosh", np.cosh), ("tanh", np.tanh), ("asinh", np.arcsinh),
                   ("acosh", np.arccosh), ("atanh", np.arctanh)):
    POINTWISE_UNARY[_name] = _via_float32(_fn)


# --------------------------------------------------------------------------------------------------------------------
# matmul / reductions -- cpu/source/dopt/cpu/math.d:70-310
# --------------------------------------------------------------------------------------------------------------------
def matmul(a, b):
    """gemm(RowMajor, NoTrans, NoTrans, alpha=1, beta=0), math.d:79-82."""
    assert a.dtype == F32 and b.dtype == F32, "Not implemented."  # math.d:86
    return np.matmul(a, b).astype(F32)


def _reduce_axes(x, axes, kind):
    # axes are reduced one after another on the full-rank shape, math.d:124-150 / 205-231
    cur = x
    for ax in axes:
        if kind == "sum":
            if cur.dtype == F32:
                cur = np.sum(cur.astype(np.float64), axis=ax, keepdims=True).astype(F32)
            else:
                cur = np.sum(cur, axis=ax, keepdims=True, dtype=np.int64).astype(np.int32)
        else:
            cur = np.max(cur, axis=ax, keepdims=True)
    out_shape = [s for i, s in enumerate(x.shape) if i not in axes]
    return cur.reshape(out_shape)


def sum_(x, axes):
    return _reduce_axes(x, list(axes), "sum")


def max_element(x, axes):
    return _reduce_axes(x, list(axes), "max")


def argmin(x, axis):
    """First minimum along `axis` (strict `<`, math.d:281); result keeps the axis with extent 1, int32."""
    return np.expand_dims(np.argmin(x, axis=axis), axis).astype(np.int32)


# --------------------------------------------------------------------------------------------------------------------
# slice / pad / transpose / repeat -- cpu/source/dopt/cpu/basic.d
# --------------------------------------------------------------------------------------------------------------------
def slice_(x, start, stop):
    return x[tuple(slice(a, b) for a, b in zip(start, stop))].copy()


def pad(x, before, after):
    return np.pad(x, list(zip(before, after)), mode="constant")


def transpose(x, order):
    # rank <= 2 only (basic.d:128); order [0,1] is the identity
    assert x.ndim <= 2, "transpose is only implemented for rank <= 2"
    return np.ascontiguousarray(np.transpose(x, order)) if x.ndim == 2 else x.copy()


def repeat(x, reps):
    return np.tile(x, reps)  # per-axis tiling, basic.d:156-204 (unit test core/ops/basic.d:336-358)


# --------------------------------------------------------------------------------------------------------------------
# nnet -- cpu/source/dopt/cpu/nnet.d + cuDNN definitions
# --------------------------------------------------------------------------------------------------------------------
def _conv_out(h, p, r, u):
    return (h + 2 * p - r) // u + 1  # judgeConvolution, core/source/dopt/core/ops/nnet.d:68-87


def _im2col(x, R, S, ph, pw, u, v, P, Q):
    """[N,C,H,W] -> [N, C*R*S, P*Q] with column order (c, r, s)."""
    N, Cc, H, W = x.shape
    xp = np.zeros((N, Cc, H + 2 * ph + u, W + 2 * pw + v), dtype=x.dtype)
    xp[:, :, ph:ph + H, pw:pw + W] = x
    cols = np.empty((N, Cc, R, S, P, Q), dtype=x.dtype)
    for r in range(R):
        for s in range(S):
            cols[:, :, r, s] = xp[:, :, r:r + u * P:u, s:s + v * Q:v][:, :, :P, :Q]
    return cols.reshape(N, Cc * R * S, P * Q)


def convolution(x, w, padding=(0, 0), stride=(1, 1)):
    """True convolution: y[n,k,p,q] = sum x[n,c,p*u-ph+r,q*v-pw+s] * w[k,c,R-1-r,S-1-s]
    (cpu/source/dopt/cpu/nnet.d:44-59 flips via jprime/iprime; cuDNN CUDNN_CONVOLUTION, cudnn7.d:87)."""
    N, Cc, H, W = x.shape
    K, C2, R, S = w.shape
    assert Cc == C2
    ph, pw = padding
    u, v = stride
    P, Q = _conv_out(H, ph, R, u), _conv_out(W, pw, S, v)
    cols = _im2col(x.astype(F32), R, S, ph, pw, u, v, P, Q)
    wf = w[:, :, ::-1, ::-1].reshape(K, Cc * R * S).astype(F32)
    y = np.matmul(wf[None], cols)  # [N, K, P*Q], sgemm
    return y.reshape(N, K, P, Q).astype(F32)


def convolution_features_grad(dy, w, features_shape, padding=(0, 0), stride=(1, 1)):
    """cudnnConvolutionBackwardData (cudnn7.d:161-204): dx = col2im( Wf^T . dy )."""
    N, Cc, H, W = features_shape
    K, C2, R, S = w.shape
    ph, pw = padding
    u, v = stride
    P, Q = dy.shape[2], dy.shape[3]
    wf = w[:, :, ::-1, ::-1].reshape(K, Cc * R * S).astype(F32)
    cols = np.matmul(wf.T[None], dy.reshape(N, K, P * Q).astype(F32))  # [N, C*R*S, P*Q]
    cols = cols.reshape(N, Cc, R, S, P, Q)
    dxp = np.zeros((N, Cc, H + 2 * ph + u, W + 2 * pw + v), dtype=np.float64)
    for r in range(R):
        for s in range(S):
            dxp[:, :, r:r + u * P:u, s:s + v * Q:v][:, :, :P, :Q] += cols[:, :, r, s]
    return dxp[:, :, ph:ph + H, pw:pw + W].astype(F32)


def convolution_filters_grad(dy, x, filters_shape, padding=(0, 0), stride=(1, 1)):
    """cudnnConvolutionBackwardFilter (cudnn7.d:206-249): dWf = sum_n dy_n . cols_n^T, then un-flip."""
    K, Cc, R, S = filters_shape
    N = x.shape[0]
    ph, pw = padding
    u, v = stride
    P, Q = dy.shape[2], dy.shape[3]
    cols = _im2col(x.astype(F32), R, S, ph, pw, u, v, P, Q)  # [N, CRS, PQ]
    g = np.zeros((K, Cc * R * S), dtype=np.float64)
    dyf = dy.reshape(N, K, P * Q).astype(F32)
    for n in range(N):
        g += np.matmul(dyf[n], cols[n].T)
    return g.reshape(K, Cc, R, S)[:, :, ::-1, ::-1].astype(F32)


def maxpool(x, dims):
    """cpu/source/dopt/cpu/nnet.d:89-127: non-overlapping window, output floor(H/dh) x floor(W/dw)."""
    dh, dw = dims
    N, Cc, H, W = x.shape
    OH, OW = H // dh, W // dw
    v = x[:, :, :OH * dh, :OW * dw].reshape(N, Cc, OH, dh, OW, dw)
    return v.max(axis=(3, 5)).astype(x.dtype)


def maxpool_grad(dy, y, x, dims, tie_all=False):
    """cudnnPoolingBackward(y, dy, x) for CUDNN_POOLING_MAX (cudnn7.d:309-333): dy goes to the element equal to the
    window maximum.  On ties cuDNN routes it to the FIRST such element in row-major window order only (measured with the
    replayed call on a B200, cuDNN 9.22: tests/test_cudnn_replay_gpu.py::test_maxpool_grad_tie_rule_is_cudnns); tie_all=True
    gives the other plausible rule (every tied element) for comparison."""
    dh, dw = dims
    N, Cc, H, W = x.shape
    OH, OW = H // dh, W // dw
    dx = np.zeros_like(x)
    xv = x[:, :, :OH * dh, :OW * dw].reshape(N, Cc, OH, dh, OW, dw)
    mask = (xv == y[:, :, :, None, :, None])
    if not tie_all:
        flat = mask.transpose(0, 1, 2, 4, 3, 5).reshape(N, Cc, OH, OW, dh * dw)
        first = np.zeros_like(flat)
        idx = flat.argmax(axis=-1)
        np.put_along_axis(first, idx[..., None], True, axis=-1)
        first &= flat
        mask = first.reshape(N, Cc, OH, OW, dh, dw).transpose(0, 1, 2, 4, 3, 5)
    g = mask.astype(x.dtype) * dy[:, :, :, None, :, None]
    dx[:, :, :OH * dh, :OW * dw] = g.reshape(N, Cc, OH * dh, OW * dw)
    return dx


def _ncv(x):
    return x.reshape(x.shape[0], x.shape[1], -1)


def softmax(x):
    """cpu/source/dopt/cpu/nnet.d:129-166: softmax over dim 1, max-subtracted."""
    v = _ncv(x).astype(F32)
    m = v.max(axis=1, keepdims=True)
    e = np.exp(v - m).astype(F32)
    s = e.sum(axis=1, keepdims=True, dtype=F32)
    return (e / s).astype(F32).reshape(x.shape)


def softmax_grad(dy, y):
    """cudnnSoftmaxBackward (cudnn7.d:373-404): dx = y * (dy - sum_c dy*y)."""
    yv, gv = _ncv(y).astype(np.float64), _ncv(dy).astype(np.float64)
    s = (yv * gv).sum(axis=1, keepdims=True)
    return (yv * (gv - s)).astype(F32).reshape(y.shape)


def relu(x):
    return np.where((x > 0) | np.isnan(x), x, F32(0)).astype(F32)  # CUDNN_PROPAGATE_NAN, cudnn7.d:415


def relu_grad(dy, y, x):
    return np.where(x > 0, dy, F32(0)).astype(F32)  # cudnnActivationBackward, cudnn7.d:439-478


def add_bias(x, b):
    shp = [1, -1] + [1] * (x.ndim - 2)
    return (x + b.reshape(shp)).astype(F32)  # cudnnAddTensor, cudnn7.d:480-512


def add_bias_grad(dy):
    """cudnnConvolutionBackwardBias (cudnn7.d:514-545), first-execution value (beta acts on a zeroed buffer)."""
    return _ncv(dy).astype(np.float64).sum(axis=(0, 2)).astype(F32)


def _bn_axes(x):
    return tuple(i for i in range(x.ndim) if i != 1)


def batch_norm_train(x, scale, bias, mean, var, momentum):
    """cudnnBatchNormalizationForwardTraining, SPATIAL (cudnn7.d:587-614).  Returns the packed rank-1 tensor
    [y | newRunningMean | newRunningVar] (core/source/dopt/core/ops/nnet.d:222-225)."""
    ax = _bn_axes(x)
    shp = [1, -1] + [1] * (x.ndim - 2)
    xd = x.astype(np.float64)
    m = xd.mean(axis=ax)
    vb = xd.var(axis=ax)  # biased: used to normalise
    M = x.size // x.shape[1]
    vu = vb * M / (M - 1) if M > 1 else vb  # unbiased: goes into the running estimate
    y = (xd - m.reshape(shp)) / np.sqrt(vb.reshape(shp) + BN_EPS) * scale.reshape(shp).astype(np.float64) \
        + bias.reshape(shp).astype(np.float64)
    f = 1.0 - float(momentum)  # exponentialAverageFactor, cudnn7.d:592
    nm = mean.astype(np.float64) * (1 - f) + m * f
    nv = var.astype(np.float64) * (1 - f) + vu * f
    return np.concatenate([y.astype(F32).ravel(), nm.astype(F32).ravel(), nv.astype(F32).ravel()])


def batch_norm_grad(dy, x, scale, out_volume=None):
    """cudnnBatchNormalizationBackward with null saved statistics (cudnn7.d:616-636): batch statistics are recomputed
    from x.  Returns packed [dx | dscale | dbias]; the reference over-allocates the buffer to vol(dy)+vol(x)+vol(scale)
    (core/ops/nnet.d:232-235) and leaves the tail zero."""
    ax = _bn_axes(x)
    shp = [1, -1] + [1] * (x.ndim - 2)
    xd, gd = x.astype(np.float64), dy.astype(np.float64)
    M = x.size // x.shape[1]
    m = xd.mean(axis=ax).reshape(shp)
    vb = xd.var(axis=ax).reshape(shp)
    istd = 1.0 / np.sqrt(vb + BN_EPS)
    xhat = (xd - m) * istd
    dbias = gd.sum(axis=ax)
    dscale = (gd * xhat).sum(axis=ax)
    g = scale.reshape(shp).astype(np.float64)
    dx = g * istd / M * (M * gd - dbias.reshape(shp) - xhat * dscale.reshape(shp))
    packed = np.concatenate([dx.astype(F32).ravel(), dscale.astype(F32).ravel(), dbias.astype(F32).ravel()])
    if out_volume is not None and out_volume > packed.size:
        packed = np.concatenate([packed, np.zeros(out_volume - packed.size, dtype=F32)])
    return packed


def batch_norm_inference(x, scale, bias, mean, var):
    shp = [1, -1] + [1] * (x.ndim - 2)
    xd = x.astype(np.float64)
    y = scale.reshape(shp).astype(np.float64) * (xd - mean.reshape(shp)) / np.sqrt(var.reshape(shp) + BN_EPS) \
        + bias.reshape(shp)
    return y.astype(F32)


# --------------------------------------------------------------------------------------------------------------------
# optimiser steps -- online/source/dopt/online/{sgd,adam,amsgrad}.d evaluated op by op in float32
# --------------------------------------------------------------------------------------------------------------------
def sgd_step(w, g, m, lr, mu, nesterov=False):
    lr, mu = F32(lr), F32(mu)
    if nesterov:  # sgd.d:46-55
        nm = (m * mu).astype(F32) - (lr * g).astype(F32)
        nw = ((w + (mu * nm).astype(F32)).astype(F32) - (lr * g).astype(F32)).astype(F32)
    else:  # sgd.d:57-64
        nm = ((m * mu).astype(F32) + (lr * g).astype(F32)).astype(F32)
        nw = (w - nm).astype(F32)
    return nw, nm


def adam_step(w, g, m, v, b1, b2, alpha, beta1, beta2, eps, vhat=None):
    """adam.d:46-66 / amsgrad.d:46-70.  Returns (w', m', v', b1', b2'[, vhat'])."""
    alpha, beta1, beta2, eps, b1, b2 = (F32(t) for t in (alpha, beta1, beta2, eps, b1, b2))
    nb1, nb2 = F32(b1 * beta1), F32(b2 * beta2)
    eta = F32(F32(alpha * np.sqrt(F32(F32(1.0) - nb2))) / F32(F32(1.0) - nb1))
    omb1, omb2 = F32(F32(1.0) - beta1), F32(F32(1.0) - beta2)
    nm = ((beta1 * m).astype(F32) + (omb1 * g).astype(F32)).astype(F32)
    nv = ((beta2 * v).astype(F32) + ((omb2 * g).astype(F32) * g).astype(F32)).astype(F32)
    nw = (w - (eta * (nm / (np.sqrt(nv).astype(F32) + eps).astype(F32)).astype(F32)).astype(F32)).astype(F32)
    if vhat is not None:
        return nw, nm, nv, nb1, nb2, np.fmax(vhat, v).astype(F32)  # max(varhat, OLD var), amsgrad.d:63-66 (F11)
    return nw, nm, nv, nb1, nb2


_UNIFORM_RNG = np.random.RandomState(0)


def uniform(shape):
    """cpu/source/dopt/cpu/random.d:18-26: `uniform01!float + float.epsilon`, i.e. (0, 1] like cuRAND's generator.  The
    reference draws from an unseeded std.random, so only the distribution is defined; this stream is seeded for tests."""
    return (_UNIFORM_RNG.random_sample(tuple(int(s) for s in shape)).astype(np.float32) + np.finfo(np.float32).eps)


# --------------------------------------------------------------------------------------------------------------------
# one entry point keyed by op type -- the shape of cpu/source/dopt/cpu/package.d's kernel registry
# --------------------------------------------------------------------------------------------------------------------
def evaluate_op(op_type, inputs, attrs=None, out_shape=None):
    a = attrs or {}
    if op_type in POINTWISE_BINARY:
        return POINTWISE_BINARY[op_type](inputs[0], inputs[1])
    if op_type in POINTWISE_UNARY:
        return POINTWISE_UNARY[op_type](inputs[0])
    if op_type == "matmul":
        return matmul(*inputs)
    if op_type == "sum":
        return sum_(inputs[0], a["axes"])
    if op_type == "maxElement":
        return max_element(inputs[0], a["axes"])
    if op_type == "argmin":
        return argmin(inputs[0], a["axis"])
    if op_type == "slice":
        return slice_(inputs[0], a["start"], a["stop"])
    if op_type == "pad":
        return pad(inputs[0], a["before"], a["after"])
    if op_type == "transpose":
        return transpose(inputs[0], a["order"])
    if op_type == "repeat":
        return repeat(inputs[0], a["repetitions"])
    if op_type == "reshape":
        return inputs[0].reshape(a["shape"])
    if op_type == "convolution":
        return convolution(inputs[0], inputs[1], a.get("padding", (0, 0)), a.get("stride", (1, 1)))
    if op_type == "convolutionFeaturesGrad":
        return convolution_features_grad(inputs[0], inputs[1], a["featuresShape"], a.get("padding", (0, 0)),
                                         a.get("stride", (1, 1)))
    if op_type == "convolutionFiltersGrad":
        return convolution_filters_grad(inputs[0], inputs[1], a["filtersShape"], a.get("padding", (0, 0)),
                                        a.get("stride", (1, 1)))
    if op_type == "maxpool":
        return maxpool(inputs[0], a["dims"])
    if op_type == "maxpoolGrad":
        return maxpool_grad(inputs[0], inputs[1], inputs[2], a["dims"])
    if op_type == "softmax":
        return softmax(inputs[0])
    if op_type == "softmaxGrad":
        return softmax_grad(inputs[0], inputs[1])
    if op_type == "relu":
        return relu(inputs[0])
    if op_type == "reluGrad":
        return relu_grad(*inputs)
    if op_type == "addBias":
        return add_bias(*inputs)
    if op_type == "addBiasGrad":
        return add_bias_grad(inputs[0])
    if op_type == "batchNormTrain":
        return batch_norm_train(*inputs, momentum=a["momentum"])
    if op_type == "batchNormGrad":
        vol = int(np.prod(out_shape)) if out_shape is not None else None
        return batch_norm_grad(*inputs, out_volume=vol)
    if op_type == "batchNormInference":
        return batch_norm_inference(*inputs)
    if op_type == "uniform":
        return uniform(a["shape"])
    raise KeyError("oracle has no kernel for '%s'" % op_type)
