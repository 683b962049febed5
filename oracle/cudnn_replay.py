"""ctypes face of oracle/_build/libcudnn_replay.so -- TEST INFRASTRUCTURE ONLY.

The replay library issues the reference CUDA backend's own cuDNN / cuBLAS call sequences (cuda/source/dopt/cuda/nnet/
cudnn7.d, cuda/source/dopt/cuda/math.d:214-247, cuda/source/dopt/cuda/basic.d:219-247) from C++; see cudnn_replay.cpp.
Functions here take and return torch CUDA tensors (fp32, contiguous) and follow the operand order of the reference
kernels' `execute(inputs, output)` (inputs[i] <-> op.deps[i]).  Only tests/ and tools/ import this module.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "_build", "libcudnn_replay.so")

_lib = None


class ReplayError(RuntimeError):
    pass


def available():
    return os.path.exists(PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise ReplayError("oracle/_build/libcudnn_replay.so is not built (run __graft_entry__.build())")
        _lib = C.CDLL(PATH)
        _lib.cudnn_replay_last_error.restype = C.c_char_p
        _lib.cudnn_replay_versions.restype = C.c_long
    return _lib


def _check(rc):
    if rc != 0:
        raise ReplayError(lib().cudnn_replay_last_error().decode("utf-8", "replace"))


def _p(t):
    return C.c_void_p(int(t.data_ptr()) if t is not None else 0)


def _ints(v):
    return (C.c_int * len(v))(*[int(i) for i in v])


def _f(t):
    import torch
    assert t.is_cuda and t.dtype == torch.float32
    return t.contiguous()


def versions():
    rt = C.c_int(0)
    v = lib().cudnn_replay_versions(C.byref(rt))
    return {"cudnn": int(v), "cudart": rt.value}


def _shape4(shape):
    s = [int(v) for v in shape][:4]
    return s + [1] * (4 - len(s))


def _ncv(shape):
    vol = 1
    for s in shape[2:]:
        vol *= int(s)
    return int(shape[0]), int(shape[1]), vol


# ---- convolution family (cudnn7.d:55-249) ---------------------------------------------------------------------------
def convolution(x, w, padding, stride, math=0):
    import torch
    x, w = _f(x), _f(w)
    N, Cc, H, W = x.shape
    K, _, R, S = w.shape
    P = (H + 2 * padding[0] - R) // stride[0] + 1
    Q = (W + 2 * padding[1] - S) // stride[1] + 1
    y = torch.zeros((N, K, P, Q), device="cuda")
    algo = C.c_int(-1)
    _check(lib().cudnn_replay_conv(0, _ints(x.shape), _ints(w.shape), _ints(y.shape), _ints(padding), _ints(stride),
                                   _p(x), _p(w), _p(y), int(math), C.byref(algo)))
    return y


def convolution_features_grad(dy, w, features_shape, padding, stride, math=0):
    import torch
    dy, w = _f(dy), _f(w)
    dx = torch.zeros(tuple(features_shape), device="cuda")
    algo = C.c_int(-1)
    _check(lib().cudnn_replay_conv(1, _ints(dx.shape), _ints(w.shape), _ints(dy.shape), _ints(padding), _ints(stride),
                                   _p(dx), _p(w), _p(dy), int(math), C.byref(algo)))
    return dx


def convolution_filters_grad(dy, x, filters_shape, padding, stride, math=0):
    import torch
    dy, x = _f(dy), _f(x)
    dw = torch.zeros(tuple(filters_shape), device="cuda")
    algo = C.c_int(-1)
    _check(lib().cudnn_replay_conv(2, _ints(x.shape), _ints(dw.shape), _ints(dy.shape), _ints(padding), _ints(stride),
                                   _p(x), _p(dw), _p(dy), int(math), C.byref(algo)))
    return dw


# ---- pooling (cudnn7.d:251-333) --------------------------------------------------------------------------------------
def maxpool(x, dims):
    import torch
    x = _f(x)
    N, Cc, H, W = x.shape
    y = torch.zeros((N, Cc, H // dims[0], W // dims[1]), device="cuda")
    _check(lib().cudnn_replay_maxpool(0, _ints(x.shape), _ints(y.shape), _ints(dims), _p(x), _p(y), _p(None), _p(None)))
    return y


def maxpool_grad(dy, y, x, dims):
    import torch
    dy, y, x = _f(dy), _f(y), _f(x)
    dx = torch.zeros_like(x)
    _check(lib().cudnn_replay_maxpool(1, _ints(x.shape), _ints(y.shape), _ints(dims), _p(x), _p(y), _p(dy), _p(dx)))
    return dx


# ---- softmax / relu / bias (cudnn7.d:335-545) ------------------------------------------------------------------------
def softmax(x):
    import torch
    x = _f(x)
    out = torch.zeros_like(x)
    n, c, v = _ncv(x.shape)
    _check(lib().cudnn_replay_softmax(0, n, c, v, _p(x), _p(None), _p(out)))
    return out


def softmax_grad(dy, y):
    import torch
    dy, y = _f(dy), _f(y)
    out = torch.zeros_like(y)
    n, c, v = _ncv(y.shape)
    _check(lib().cudnn_replay_softmax(1, n, c, v, _p(y), _p(dy), _p(out)))
    return out


def relu(x):
    import torch
    x = _f(x)
    out = torch.zeros_like(x)
    n, c, v = _ncv(x.shape)
    _check(lib().cudnn_replay_relu(0, n, c, v, _p(x), _p(None), _p(None), _p(out)))
    return out


def relu_grad(dy, y, x):
    import torch
    dy, y, x = _f(dy), _f(y), _f(x)
    out = torch.zeros_like(x)
    n, c, v = _ncv(x.shape)
    _check(lib().cudnn_replay_relu(1, n, c, v, _p(x), _p(y), _p(dy), _p(out)))
    return out


def add_bias(x, b):
    import torch
    x, b = _f(x), _f(b)
    out = torch.zeros_like(x)
    n, c, v = _ncv(x.shape)
    _check(lib().cudnn_replay_add_bias(n, c, v, _p(x), _p(b), _p(out)))
    return out


def add_bias_grad(dy, executions=1):
    """`executions` > 1 shows the reference's beta = 1 accumulation across plan executions."""
    import torch
    dy = _f(dy)
    n, c, v = _ncv(dy.shape)
    out = torch.zeros((c,), device="cuda")          # the plan buffer: zeroed at creation only (package.d:152)
    for _ in range(executions):
        _check(lib().cudnn_replay_add_bias_grad(n, c, v, _p(dy), _p(out)))
    return out


# ---- batch norm (cudnn7.d:547-654) -----------------------------------------------------------------------------------
def batch_norm_train(x, scale, bias, mean, var, momentum):
    import torch
    x, scale, bias, mean, var = _f(x), _f(scale), _f(bias), _f(mean), _f(var)
    c = int(x.shape[1])
    packed = torch.zeros((x.numel() + 2 * c,), device="cuda")
    _check(lib().cudnn_replay_bn_train(_ints(_shape4(x.shape)), C.c_double(momentum), _p(x), _p(scale), _p(bias),
                                       _p(mean), _p(var), _p(packed)))
    return packed


def batch_norm_grad(dy, x, scale):
    import torch
    dy, x, scale = _f(dy), _f(x), _f(scale)
    c = int(x.shape[1])
    packed = torch.zeros((2 * x.numel() + c,), device="cuda")   # over-allocated like core/ops/nnet.d:232-235
    _check(lib().cudnn_replay_bn_grad(_ints(_shape4(x.shape)), _p(dy), _p(x), _p(scale), _p(packed)))
    return packed


def batch_norm_inference(x, scale, bias, mean, var):
    import torch
    x, scale, bias, mean, var = _f(x), _f(scale), _f(bias), _f(mean), _f(var)
    y = torch.zeros_like(x)
    _check(lib().cudnn_replay_bn_inference(x.dim(), _ints(_shape4(x.shape)), _p(x), _p(scale), _p(bias), _p(mean),
                                           _p(var), _p(y)))
    return y


# ---- cuBLAS (math.d:214-247, basic.d:219-247) ------------------------------------------------------------------------
def matmul(a, b):
    import torch
    a, b = _f(a), _f(b)
    M, K = a.shape
    N = b.shape[1]
    c = torch.zeros((M, N), device="cuda")
    _check(lib().cudnn_replay_matmul(int(M), int(K), int(N), _p(a), _p(b), _p(c)))
    return c


def transpose(a):
    import torch
    a = _f(a)
    rows_out, cols_out = int(a.shape[1]), int(a.shape[0])
    c = torch.zeros((rows_out, cols_out), device="cuda")
    _check(lib().cudnn_replay_transpose(rows_out, cols_out, _p(a), _p(c)))
    return c
