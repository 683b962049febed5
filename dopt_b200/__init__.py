"""dopt_b200 -- B200 (sm_100a) kernels for dopt's CUDA training hot path, behind dopt's kernel-plugin API.

This package is only the thin Python face of two native libraries built from `dopt_b200/csrc` and `dopt_b200/host`:

  lib/libdopt_b200.so   hand-written CUDA kernels + the C ABI of `include/dopt_b200.h`  (the product)
  lib/libdopt_host.so   C++ mirror of dopt's host side (Operation graph, autodiff, nnet layers, dopt.online updaters,
                        CUDAPlan) that calls ONLY that C ABI -- it stands in for the D host because no D compiler
                        exists in this environment (see INTEGRATION.md for the D glue a maintainer would add).

PyTorch is used by tests and bench.py for device memory, streams and torch.distributed; it is not on the compute path.
There is no CPU fallback anywhere: importing works without a GPU (so the ABI can be inspected), every compute call fails
loudly without one.
"""
import ctypes as C

from . import _lib

__all__ = ["lib", "DoptError", "check", "make_op", "CUDAKernel", "run_op", "list_operations", "FLOAT32", "INT32",
           "MATH_DEFAULT", "MATH_FP32", "MATH_BF16", "image_transform", "one_hot", "jitter_sample"]

FLOAT32, INT32 = _lib.FLOAT32, _lib.INT32
MATH_DEFAULT, MATH_FP32, MATH_BF16 = _lib.MATH_DEFAULT, _lib.MATH_FP32, _lib.MATH_BF16

lib = _lib.load()


class DoptError(RuntimeError):
    """Raised when a C ABI call returns non-zero (the D glue throws via `enforce` at the same place)."""


def check(rc):
    if rc != 0:
        raise DoptError(lib.dopt_b200_last_error().decode("utf-8", "replace"))


def list_operations():
    """== listCUDAOperations() minus the two pseudo ops (cuda/source/dopt/cuda/package.d:503-506)."""
    p = lib.dopt_b200_list_operations()
    out, cur, i = [], b"", 0
    while True:
        ch = p[i]
        i += 1
        if ch == b"\0":
            if not cur:
                break
            out.append(cur.decode())
            cur = b""
        else:
            cur += ch
    return out


def _fill_tensor(t, shape, dtype):
    shape = tuple(int(s) for s in shape)
    if len(shape) > _lib.MAX_RANK:
        raise DoptError("rank > %d" % _lib.MAX_RANK)
    t.dtype = dtype
    t.rank = len(shape)
    for i, s in enumerate(shape):
        t.shape[i] = s


_LIST_ATTRS = ("padding", "stride", "start", "stop", "before", "after", "repetitions", "order")


def make_op(op_type, in_shapes, out_shape, in_dtypes=None, out_dtype=FLOAT32, math=MATH_DEFAULT, **attrs):
    """Builds the POD description of one graph node (dopt `Operation`: opType, deps' types, outputType, attributes)."""
    op = _lib.Op()
    op._keep = op_type.encode()
    op.op_type = op._keep
    if len(in_shapes) > _lib.MAX_INPUTS:
        raise DoptError("too many operands")
    op.n_inputs = len(in_shapes)
    in_dtypes = in_dtypes or [FLOAT32] * len(in_shapes)
    for i, (s, d) in enumerate(zip(in_shapes, in_dtypes)):
        _fill_tensor(op.inputs[i], s, d)
    _fill_tensor(op.output, out_shape, out_dtype)
    op.stride[0] = op.stride[1] = 1
    op.math = math
    for k, v in attrs.items():
        if k in _LIST_ATTRS:
            arr = getattr(op, k)
            for i, x in enumerate(v):
                arr[i] = int(x)
        elif k == "dims":
            op.pool_dims[0], op.pool_dims[1] = int(v[0]), int(v[1])
        elif k == "axes":
            op.n_axes = len(v)
            for i, x in enumerate(v):
                op.axes[i] = int(x)
        elif k == "axis":
            op.axis = int(v)
        elif k == "momentum":
            op.momentum = float(v)
        elif k == "seed":
            op.seed = int(v)
        elif k in ("featuresShape", "filtersShape", "shape"):
            pass  # carried by the output type
        else:
            raise DoptError("unknown attribute '%s'" % k)
    return op


class CUDAKernel(object):
    """One constructed kernel: `CUDAKernel k = ctor(op)` then `k.execute(inputs, output)` every step
    (cuda/source/dopt/cuda/package.d:68-79,284-288,412)."""

    def __init__(self, op):
        self._h = C.c_void_p()
        self._op = op
        check(lib.dopt_b200_kernel_create(C.byref(op), C.byref(self._h)))

    def execute(self, inputs, output, stream=None):
        n = len(inputs)
        arr = (C.c_void_p * max(n, 1))(*[int(t.data_ptr()) for t in inputs])
        check(lib.dopt_b200_kernel_execute(self._h, arr, n, C.c_void_p(int(output.data_ptr())),
                                           C.c_void_p(stream or 0)))

    def close(self):
        if self._h:
            lib.dopt_b200_kernel_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run_op(op_type, inputs, out_shape, out_dtype=FLOAT32, math=MATH_DEFAULT, **attrs):
    """Convenience for tests: torch CUDA tensors in, a fresh torch CUDA tensor out, through the C ABI."""
    import torch

    tdt = {FLOAT32: torch.float32, INT32: torch.int32}
    rev = {torch.float32: FLOAT32, torch.int32: INT32}
    ins = [t.contiguous() for t in inputs]
    op = make_op(op_type, [tuple(t.shape) for t in ins], out_shape, [rev[t.dtype] for t in ins], out_dtype, math,
                 **attrs)
    out = torch.zeros(tuple(out_shape), dtype=tdt[out_dtype], device="cuda")
    k = CUDAKernel(op)
    k.execute(ins, out, torch.cuda.current_stream().cuda_stream)
    k.close()
    return out


# ---- on-device input pipeline (include/dopt_b200.h, "on-device input pipeline") ------------------------------------------
def image_transform(src, jitter_x=0, jitter_y=0, per_image=None, stream=None):
    """src: torch CUDA tensor [N, C, H, W], uint8 (normalised x/128-1 on the fly) or float32.  per_image: None or an int32
    CUDA tensor [N, 4] = (x_off, y_off, flip_x, flip_y).  Returns the float32 NCHW batch the plan reads."""
    import torch

    src = src.contiguous()
    n, c, h, w = (int(v) for v in src.shape)
    dst = torch.empty((n, c, h, w), dtype=torch.float32, device=src.device)
    fn = {torch.uint8: lib.dopt_b200_image_transform_u8, torch.float32: lib.dopt_b200_image_transform_f32}[src.dtype]
    if per_image is not None:
        per_image = per_image.contiguous()
        assert per_image.dtype == torch.int32 and tuple(per_image.shape) == (n, 4)
    check(fn(src.data_ptr(), dst.data_ptr(), n, c, h, w, int(jitter_x), int(jitter_y),
             per_image.data_ptr() if per_image is not None else None,
             stream if stream is not None else torch.cuda.current_stream().cuda_stream))
    return dst


def one_hot(labels, classes, stream=None):
    import torch

    labels = labels.contiguous()
    assert labels.dtype == torch.uint8
    dst = torch.empty((labels.numel(), int(classes)), dtype=torch.float32, device=labels.device)
    check(lib.dopt_b200_one_hot_u8(labels.data_ptr(), dst.data_ptr(), labels.numel(), int(classes),
                                   stream if stream is not None else torch.cuda.current_stream().cuda_stream))
    return dst


def jitter_sample(n, jitter_x, jitter_y, flip_x, flip_y, seed, call=0, stream=None):
    import torch

    out = torch.empty((int(n), 4), dtype=torch.int32, device="cuda")
    check(lib.dopt_b200_jitter_sample(out.data_ptr(), int(n), int(jitter_x), int(jitter_y), int(bool(flip_x)),
                                      int(bool(flip_y)), int(seed), int(call),
                                      stream if stream is not None else torch.cuda.current_stream().cuda_stream))
    return out
