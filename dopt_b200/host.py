"""Python face of libdopt_host.so -- the C++ mirror of dopt's host side (graph, autodiff, nnet layers, dopt.online).

Used by tests and bench.py to build dopt graphs the way a D program would (same function names and lowering), export
them for the CPU oracle, and run plans / updaters on the GPU.  Nothing numerical happens in Python.
"""
import ctypes as C
import os

import numpy as np

from . import _lib

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "lib", "libdopt_host.so")

if not os.path.exists(HOST_LIB_PATH):
    raise ImportError("libdopt_host.so not found at %s -- run __graft_entry__.build()" % HOST_LIB_PATH)
_lib.load()  # libdopt_b200.so first (RTLD_GLOBAL) so the host library resolves the C ABI against it
_h = C.CDLL(HOST_LIB_PATH)

vp, i32p, i64p = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int64)
_h.dh_last_error.restype = C.c_char_p
_h.dh_init_error.restype = C.c_char_p
_h.dh_export.restype = C.c_char_p
_h.dh_updater_profile.restype = C.c_char_p
_h.dh_value_device_ptr.restype = vp
_h.dh_op_serial.restype = C.c_int64
_h.dh_set_stream.argtypes = [vp]
_h.dh_seed.argtypes = [C.c_uint64]
_h.dh_variable.argtypes = [C.c_int, i64p, C.c_int, vp]
_h.dh_constant.argtypes = [C.c_int, i64p, C.c_int, vp]
_h.dh_create.argtypes = [C.c_char_p, i32p, C.c_int, C.c_int, C.POINTER(C.c_char_p), i32p, i64p, i32p,
                         C.POINTER(C.c_double)]
_h.dh_binary_scalar.argtypes = [C.c_int, C.c_int, C.c_float, C.c_int]
_h.dh_repeat_n.argtypes = [C.c_int, C.c_int64]
_h.dh_sum.argtypes = [C.c_int, i64p, C.c_int]
_h.dh_max_element.argtypes = [C.c_int, i64p, C.c_int]
_h.dh_batch_norm_train.argtypes = [C.c_int] * 5 + [C.c_double, i32p]
_h.dh_convolution_transpose.argtypes = [C.c_int, C.c_int, i64p, i64p]
_h.dh_grad.argtypes = [C.c_int, i32p, C.c_int, i32p]
_h.dh_op_shape.argtypes = [C.c_int, i64p]
_h.dh_get_value.argtypes = [C.c_int, vp, C.c_size_t]
_h.dh_set_value.argtypes = [C.c_int, vp, C.c_size_t]
_h.dh_export.argtypes = [i32p, C.c_int]
_h.dh_conv2d.argtypes = [C.c_int, C.c_int64, i64p, i64p, i64p, C.c_float, C.c_int]
_h.dh_dense.argtypes = [C.c_int, C.c_int64, C.c_float, C.c_int]
_h.dh_batch_norm.argtypes = [C.c_int, C.c_float]
_h.dh_conv2d_reg.argtypes = [C.c_int, C.c_int64, i64p, i64p, i64p, C.c_float, C.c_int, C.c_float, C.c_float]
_h.dh_dense_reg.argtypes = [C.c_int, C.c_int64, C.c_float, C.c_int, C.c_float, C.c_float]
_h.dh_batch_norm_reg.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float]
_h.dh_wide_resnet_reg.argtypes = [C.c_int, C.c_int64, C.c_int64, i64p, C.c_float, C.c_int, C.c_float, C.c_float, C.c_float,
                                  C.c_float]
_h.dh_vgg_reg.argtypes = [C.c_int, C.c_int, i64p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float]
_h.dh_max_pool.argtypes = [C.c_int, i64p]
_h.dh_dropout.argtypes = [C.c_int, C.c_float]
_h.dh_wide_resnet.argtypes = [C.c_int, C.c_int64, C.c_int64, i64p, C.c_float]
_h.dh_vgg19.argtypes = [C.c_int, i64p, C.c_int, C.c_int]
_h.dh_network.argtypes = [i32p, C.c_int, i32p, C.c_int]
_h.dh_network_params.argtypes = [C.c_int, i32p, C.c_int]
_h.dh_network_save.argtypes = [C.c_int, C.c_char_p]
_h.dh_network_load.argtypes = [C.c_int, C.c_char_p]
_h.dh_compile.argtypes = [i32p, C.c_int, C.c_int]
_h.dh_plan_execute.argtypes = [C.c_int, i32p, C.POINTER(vp), C.POINTER(C.c_size_t), C.c_int, C.POINTER(vp)]
_h.dh_plan_stats.argtypes = [C.c_int, i64p, i64p, i64p]
_h.dh_updater.argtypes = [C.c_int, i32p, C.c_int, C.c_int, i32p, C.c_int, i32p]
_h.dh_updater_step.argtypes = [C.c_int, i32p, C.POINTER(vp), C.POINTER(C.c_size_t), C.c_int, C.POINTER(vp)]
_h.dh_updater_step_device.argtypes = [C.c_int, i32p, C.POINTER(vp), C.c_int]
_h.dh_updater_plan_outputs.argtypes = [C.c_int, i32p, i32p, C.c_int]
_h.dh_updater_stats.argtypes = [C.c_int, i64p, i64p, i64p]
_h.dh_updater_replay_class.argtypes = [C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_double), i64p]
_h.dh_init_data_parallel.argtypes = [C.c_int, C.c_int, vp]
_h.dh_matrix_norm.argtypes = [C.c_int, C.c_int]
_h.dh_conv_params_norm.argtypes = [C.c_int, i64p, i64p, i64p, C.c_int]
_h.dh_max_norm.argtypes = [C.c_int, C.c_int, C.c_int]
_h.dh_updater_save_state.argtypes = [C.c_int, C.c_char_p]
_h.dh_updater_load_state.argtypes = [C.c_int, C.c_char_p]
_h.dh_updater_state_header_bytes.argtypes = [C.c_int]
_h.dh_updater_state_header_bytes.restype = C.c_int64


class HostError(RuntimeError):
    pass


def _ck(rc):
    if rc is None or rc < 0:
        raise HostError(_h.dh_last_error().decode("utf-8", "replace"))
    return rc


def _i64(seq):
    seq = [int(x) for x in seq]
    return (C.c_int64 * max(len(seq), 1))(*seq)


def _i32(seq):
    seq = [int(x) for x in seq]
    return (C.c_int * max(len(seq), 1))(*seq)


_NP = {0: np.float32, 1: np.int32}
_backend = None


def init():
    """dopt.core + dopt.cuda module constructors.  Returns True when the CUDA backend is up."""
    global _backend
    if _backend is None:
        _backend = _ck(_h.dh_init()) == 1
    return _backend


def init_error():
    return _h.dh_init_error().decode()


def reset():
    _h.dh_reset()


def seed(s):
    _h.dh_seed(int(s))


def set_plan_flags(flags):
    _h.dh_set_plan_flags(int(flags))


def plan_flags():
    return int(_h.dh_plan_flags())


def set_math(m):
    _h.dh_set_math(int(m))


def set_stream(s):
    _h.dh_set_stream(C.c_void_p(int(s)))


def set_data_parallel_world(n):
    _h.dh_set_data_parallel_world(int(n))


def init_data_parallel(rank, world, unique_id):
    buf = C.create_string_buffer(bytes(unique_id), 128)
    _ck(_h.dh_init_data_parallel(rank, world, buf))


class Op(object):
    """Handle of a dopt `Operation`."""

    __slots__ = ("h",)

    def __init__(self, h):
        self.h = _ck(h)

    @property
    def shape(self):
        buf = (C.c_int64 * 8)()
        n = _ck(_h.dh_op_shape(self.h, buf))
        return tuple(int(buf[i]) for i in range(n))

    @property
    def dtype(self):
        return _NP[_ck(_h.dh_op_dtype(self.h))]

    @property
    def volume(self):
        return int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1

    @property
    def serial(self):
        return int(_h.dh_op_serial(self.h))

    def get(self):
        """`op.value.get!T` -- the variable's / constant's current contents."""
        out = np.empty(self.shape, dtype=self.dtype)
        _ck(_h.dh_get_value(self.h, out.ctypes.data_as(vp), out.nbytes))
        return out

    def set(self, arr):
        arr = np.ascontiguousarray(arr, dtype=self.dtype)
        _ck(_h.dh_set_value(self.h, arr.ctypes.data_as(vp), arr.nbytes))

    def device_ptr(self):
        return _h.dh_value_device_ptr(self.h)

    # D operator overloads (ops/package.d:94-178)
    def _bin(self, ch, other, left=False):
        if isinstance(other, Op):
            return Op(_h.dh_binary(ord(ch), other.h, self.h) if left else _h.dh_binary(ord(ch), self.h, other.h))
        return Op(_h.dh_binary_scalar(ord(ch), self.h, float(other), 1 if left else 0))

    def __add__(self, o): return self._bin("+", o)
    def __radd__(self, o): return self._bin("+", o, True)
    def __sub__(self, o): return self._bin("-", o)
    def __rsub__(self, o): return self._bin("-", o, True)
    def __mul__(self, o): return self._bin("*", o)
    def __rmul__(self, o): return self._bin("*", o, True)
    def __truediv__(self, o): return self._bin("/", o)
    def __rtruediv__(self, o): return self._bin("/", o, True)
    def __neg__(self): return create("neg", [self])


def _data_ptr(arr):
    return arr.ctypes.data_as(vp) if arr is not None else None


def variable(shape, data=None, dtype=np.float32):
    arr = None if data is None else np.ascontiguousarray(data, dtype=dtype).reshape(shape)
    return Op(_h.dh_variable(0 if dtype == np.float32 else 1, _i64(shape), len(shape), _data_ptr(arr)))


def float32(shape=(), data=None):
    return variable(tuple(shape), data, np.float32)


def int32(shape=(), data=None):
    return variable(tuple(shape), data, np.int32)


def constant(shape, data, dtype=np.float32):
    arr = np.ascontiguousarray(data, dtype=dtype).reshape(shape)
    return Op(_h.dh_constant(0 if dtype == np.float32 else 1, _i64(shape), len(shape), _data_ptr(arr)))


def create(op_type, deps, **attrs):
    """createOperation(opType, deps, attributes)."""
    names, kinds, ivals, lens, dvals = [], [], [], [], []
    for k, v in attrs.items():
        names.append(k.encode())
        if isinstance(v, (list, tuple)):
            kinds.append(1)
            ivals.extend(int(x) for x in v)
            lens.append(len(v))
        elif isinstance(v, float):
            kinds.append(3)
            dvals.append(v)
            lens.append(1)
        else:
            kinds.append(2)
            ivals.append(int(v))
            lens.append(1)
    n = len(names)
    arr_names = (C.c_char_p * max(n, 1))(*names)
    return Op(_h.dh_create(op_type.encode(), _i32([d.h for d in deps]), len(deps), n, arr_names, _i32(kinds),
                           _i64(ivals), _i32(lens), (C.c_double * max(len(dvals), 1))(*dvals)))


# the named constructors of dopt.core.ops
def slice_(x, start, stop): return create("slice", [x], start=list(start), stop=list(stop))
def pad(x, before, after): return create("pad", [x], before=list(before), after=list(after))
def reshape(x, shape): return create("reshape", [x], shape=list(shape))
def transpose(x, order): return create("transpose", [x], order=list(order))
def repeat(x, reps): return Op(_h.dh_repeat_n(x.h, reps)) if isinstance(reps, int) else create("repeat", [x], repetitions=list(reps))
def matmul(a, b): return create("matmul", [a, b])
def sum_(x, axes=()): return Op(_h.dh_sum(x.h, _i64(axes), len(axes)))
def max_element(x, axes=()): return Op(_h.dh_max_element(x.h, _i64(axes), len(axes)))
def argmin(x, axis): return create("argmin", [x], axis=int(axis))
def convolution(x, w, padding=(0, 0), stride=(1, 1)): return create("convolution", [x, w], padding=list(padding), stride=list(stride))
def convolution_transpose(x, w, padding=(0, 0), stride=(1, 1)): return Op(_h.dh_convolution_transpose(x.h, w.h, _i64(padding), _i64(stride)))
def maxpool(x, dims): return create("maxpool", [x], dims=list(dims))
def softmax(x): return create("softmax", [x])
def relu(x): return create("relu", [x])
def add_bias(x, b): return create("addBias", [x, b])


def batch_norm_train(x, scale, bias, mean, var, momentum=0.9):
    out = (C.c_int * 3)()
    _ck(_h.dh_batch_norm_train(x.h, scale.h, bias.h, mean.h, var.h, float(momentum), out))
    return [Op(out[i]) for i in range(3)]


def batch_norm_inference(x, scale, bias, mean, var): return create("batchNormInference", [x, scale, bias, mean, var])
def unary(name, x): return create(name, [x])
def binary(name, a, b): return create(name, [a, b])


def grad(objective, wrt):
    out = (C.c_int * len(wrt))()
    _ck(_h.dh_grad(objective.h, _i32([w.h for w in wrt]), len(wrt), out))
    return [Op(out[i]) for i in range(len(wrt))]


def _p_code(p): return 0 if p == float("inf") else int(p)
def matrix_norm(param, p=2): return Op(_h.dh_matrix_norm(param.h, _p_code(p)))                    # nnet/lipschitz.d:43-97
def conv_params_norm(param, in_shape, stride=(1, 1), padding=(0, 0), p=2):                        # nnet/lipschitz.d:111-147
    return Op(_h.dh_conv_params_norm(param.h, _i64(in_shape), _i64(stride), _i64(padding), _p_code(p)))
def max_norm(param, norm, maxval): return Op(_h.dh_max_norm(param.h, norm.h, maxval.h))           # nnet/lipschitz.d:162-165


def cross_entropy(hyp, truth): return Op(_h.dh_cross_entropy(hyp.h, truth.h))
def squared_error(hyp, truth): return Op(_h.dh_squared_error(hyp.h, truth.h))


def export(outputs):
    """Toposorted node list of the graph reaching `outputs` (for the CPU oracle)."""
    txt = _h.dh_export(_i32([o.h for o in outputs]), len(outputs))
    if txt is None:
        raise HostError(_h.dh_last_error().decode())
    nodes = []
    for line in txt.decode().splitlines():
        serial, typ, dt, shape, deps, attrs, handle = line.split("|")
        a = {}
        if attrs:
            for item in attrs.split(";"):
                name, kind, val = item.split(":")
                if kind == "1":
                    a[name] = [int(x) for x in val.split(",")] if val else []
                elif kind == "2":
                    a[name] = int(val)
                else:
                    a[name] = float(val)
        nodes.append({"id": int(serial), "type": typ, "dtype": _NP[int(dt)],
                      "shape": tuple(int(x) for x in shape.split(",")) if shape else (),
                      "deps": [int(x) for x in deps.split(",")] if deps else [], "attrs": a, "op": Op(int(handle))})
    return nodes


# ---- layers / networks (dopt.nnet) --------------------------------------------------------------------------------------
class Layer(object):
    def __init__(self, h):
        self.h = _ck(h)

    @property
    def output(self): return Op(_h.dh_layer_output(self.h, 0))
    @property
    def train_output(self): return Op(_h.dh_layer_output(self.h, 1))
    def conv2d(self, channels, fdims, padding=(0, 0), stride=(1, 1), weight_decay=0.0, use_bias=True,
               maxgain=float("inf"), spectral_decay=0.0):
        if maxgain == float("inf") and spectral_decay == 0.0:
            return Layer(_h.dh_conv2d(self.h, channels, _i64(fdims), _i64(padding), _i64(stride), weight_decay, int(use_bias)))
        return Layer(_h.dh_conv2d_reg(self.h, channels, _i64(fdims), _i64(padding), _i64(stride), weight_decay, int(use_bias),
                                      maxgain, spectral_decay))

    def dense(self, outputs, weight_decay=0.0, use_bias=True, maxgain=float("inf"), spectral_decay=0.0):
        if maxgain == float("inf") and spectral_decay == 0.0:
            return Layer(_h.dh_dense(self.h, outputs, weight_decay, int(use_bias)))
        return Layer(_h.dh_dense_reg(self.h, outputs, weight_decay, int(use_bias), maxgain, spectral_decay))

    def batch_norm(self, momentum=0.9, maxgain=float("inf"), lipschitz=float("inf")):
        if maxgain == float("inf") and lipschitz == float("inf"):
            return Layer(_h.dh_batch_norm(self.h, momentum))
        return Layer(_h.dh_batch_norm_reg(self.h, momentum, maxgain, lipschitz))
    def relu(self): return Layer(_h.dh_relu(self.h))
    def max_pool(self, dims): return Layer(_h.dh_max_pool(self.h, _i64(dims)))
    def dropout(self, drop_prob): return Layer(_h.dh_dropout(self.h, float(drop_prob)))
    def softmax(self): return Layer(_h.dh_softmax(self.h))


def data_source(var): return Layer(_h.dh_data_source(var.h))
def wide_resnet(features, depth, width, stride=(1, 2, 2), weight_decay=1e-4, dropout=False, maxgain_norm=float("nan"),
                lipschitz_norm=float("nan"), max_norm=float("inf"), spectral_decay=0.0):
    """wideResNet(features, depth, width, WRNOptions) -- nnet/models/wrn.d:56-102; the regulariser fields default to off."""
    if not dropout and maxgain_norm != maxgain_norm and lipschitz_norm != lipschitz_norm and spectral_decay == 0.0:
        return Layer(_h.dh_wide_resnet(features.h, depth, width, _i64(stride), weight_decay))
    return Layer(_h.dh_wide_resnet_reg(features.h, depth, width, _i64(stride), weight_decay, int(dropout), maxgain_norm,
                                       lipschitz_norm, max_norm, spectral_decay))
def vgg19(features, dense_sizes=(4096, 4096), batchnorm=False):
    return Layer(_h.dh_vgg19(features.h, _i64(dense_sizes), len(dense_sizes), int(batchnorm)))
def vgg(features, layers=19, dense_sizes=(4096, 4096), batchnorm=False, dropout=False, maxgain_norm=float("nan"),
        lipschitz_norm=float("nan"), max_norm=float("inf"), spectral_decay=0.0):
    """vgg16 / vgg19 with the full VGGOptions (nnet/models/vgg.d:12-84)."""
    return Layer(_h.dh_vgg_reg(features.h, layers, _i64(dense_sizes), len(dense_sizes), int(batchnorm), int(dropout),
                               maxgain_norm, lipschitz_norm, max_norm, spectral_decay))


class Network(object):
    """DAGNetwork (nnet/networks.d)."""

    def __init__(self, inputs, output_layers):
        self.h = _ck(_h.dh_network(_i32([i.h for i in inputs]), len(inputs), _i32([l.h for l in output_layers]),
                                   len(output_layers)))

    @property
    def param_loss(self): return Op(_h.dh_network_param_loss(self.h))
    @property
    def params(self):
        n = _ck(_h.dh_network_params(self.h, (C.c_int * 1)(), 0))
        buf = (C.c_int * n)()
        _ck(_h.dh_network_params(self.h, buf, n))
        return [Op(buf[i]) for i in range(n)]
    def save(self, path): _ck(_h.dh_network_save(self.h, path.encode()))
    def load(self, path): _ck(_h.dh_network_load(self.h, path.encode()))


# ---- plans / updaters -----------------------------------------------------------------------------------------------------
def _args(args):
    ops = list(args.keys())
    arrs = [np.ascontiguousarray(args[o], dtype=o.dtype) for o in ops]
    ptrs = (vp * max(len(ops), 1))(*[a.ctypes.data for a in arrs])
    sizes = (C.c_size_t * max(len(ops), 1))(*[a.nbytes for a in arrs])
    return ops, arrs, ptrs, sizes


class Plan(object):
    """compile(outputs) (core/package.d:123).  kind=1 builds the reference-style node-by-node CUDAPlan instead of the
    whole-graph B200Plan."""

    def __init__(self, outputs, kind=0):
        self.outputs = list(outputs)
        self.h = _ck(_h.dh_compile(_i32([o.h for o in outputs]), len(outputs), kind))

    def execute(self, args=None):
        ops, arrs, ptrs, sizes = _args(args or {})
        outs = [np.empty(o.shape, dtype=o.dtype) for o in self.outputs]
        optrs = (vp * max(len(outs), 1))(*[o.ctypes.data for o in outs])
        _ck(_h.dh_plan_execute(self.h, _i32([o.h for o in ops]), ptrs, sizes, len(ops), optrs))
        return outs

    def stats(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        _ck(_h.dh_plan_stats(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return {"launches": a.value, "device_bytes": b.value, "lowered_nodes": c.value}


def evaluate(outputs, args=None):
    return Plan(outputs).execute(args)


SGD, SGD_NESTEROV, ADAM, AMSGRAD = 0, 1, 2, 3


class Updater(object):
    """dopt.online.sgd / adam / amsgrad: returns the closure; `step(args)` is `updater(args)` + `.get!float`."""

    def __init__(self, kind, outputs, network=None, wrt=None, hyper=None):
        self.outputs = list(outputs)
        hy = _i32([h.h if h is not None else -1 for h in (hyper or [None] * 4)] + [-1] * 4)
        w = _i32([o.h for o in (wrt or [])])
        self.h = _ck(_h.dh_updater(kind, _i32([o.h for o in outputs]), len(outputs),
                                   network.h if network is not None else -1, w, len(wrt or []), hy))

    def step(self, args=None, fetch=True):
        ops, arrs, ptrs, sizes = _args(args or {})
        outs = [np.empty(o.shape, dtype=o.dtype) for o in self.outputs] if fetch else []
        optrs = (vp * max(len(self.outputs), 1))(*([o.ctypes.data for o in outs] if fetch else [None] * len(self.outputs)))
        _ck(_h.dh_updater_step(self.h, _i32([o.h for o in ops]), ptrs, sizes, len(ops), optrs))
        return outs

    def step_raw(self, op_handles, host_ptrs, nbytes, out_ptrs):
        """bench path: pre-marshalled ctypes arrays (pinned host buffers in, host pointers out)."""
        _ck(_h.dh_updater_step(self.h, op_handles, host_ptrs, nbytes, len(op_handles), out_ptrs))

    def step_device(self, op_handles, dev_ptrs):
        _ck(_h.dh_updater_step_device(self.h, op_handles, dev_ptrs, len(op_handles)))

    def plan_outputs(self):
        n = _ck(_h.dh_updater_plan_outputs(self.h, (C.c_int * 1)(), (C.c_int * 1)(), 0))
        a, b = (C.c_int * n)(), (C.c_int * n)()
        _ck(_h.dh_updater_plan_outputs(self.h, a, b, n))
        return [Op(a[i]) for i in range(n)], [Op(b[i]) if b[i] >= 0 else None for i in range(n)]

    def save_state(self, path):
        """Training checkpoint: parameters + optimiser state, header + raw fp32 (online.hpp)."""
        _ck(_h.dh_updater_save_state(self.h, path.encode()))

    def load_state(self, path): _ck(_h.dh_updater_load_state(self.h, path.encode()))

    def state_header_bytes(self): return _ck(_h.dh_updater_state_header_bytes(self.h))

    def stats(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        _ck(_h.dh_updater_stats(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return {"launches": a.value, "device_bytes": b.value, "lowered_nodes": c.value}

    def replay_class(self, op_types, reps=3):
        """(microseconds, launches) of one repetition of the kernels the profiler books under `op_types` (a comma-separated
        string), re-issued back to back with the operands of the last step -- see dopt_b200_plan_replay_class.  Leaves the
        updater's state garbage: call it last."""
        us, n = C.c_double(), C.c_int64()
        _ck(_h.dh_updater_replay_class(self.h, op_types.encode(), int(reps), C.byref(us), C.byref(n)))
        return us.value, n.value

    def profile(self, enable):
        t = _h.dh_updater_profile(self.h, int(enable))
        if t is None:
            raise HostError(_h.dh_last_error().decode())
        return dict((k, int(v)) for k, v in (l.split("=") for l in t.decode().splitlines() if l))
