"""ctypes binding of libdopt_b200.so -- the C ABI declared in include/dopt_b200.h.

Nothing in here computes anything: it marshals POD structs and raw device pointers, exactly what the D glue in
INTEGRATION.md does with `extern(C)` declarations.  If the shared library is missing the import fails loudly; there is
no Python or CPU fallback for any operation.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdopt_b200.so")

MAX_RANK = 8
MAX_INPUTS = 8

FLOAT32, INT32 = 0, 1
MATH_DEFAULT, MATH_FP32, MATH_BF16 = 0, 1, 2
PLAN_FUSE, PLAN_CUDA_GRAPH, PLAN_BF16_INTERIOR = 1, 2, 4


class Tensor(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("rank", C.c_int32), ("shape", C.c_int64 * MAX_RANK)]


class Op(C.Structure):
    _fields_ = [
        ("op_type", C.c_char_p),
        ("n_inputs", C.c_int32),
        ("inputs", Tensor * MAX_INPUTS),
        ("output", Tensor),
        ("padding", C.c_int64 * 2),
        ("stride", C.c_int64 * 2),
        ("pool_dims", C.c_int64 * 2),
        ("start", C.c_int64 * MAX_RANK),
        ("stop", C.c_int64 * MAX_RANK),
        ("before", C.c_int64 * MAX_RANK),
        ("after", C.c_int64 * MAX_RANK),
        ("repetitions", C.c_int64 * MAX_RANK),
        ("order", C.c_int64 * MAX_RANK),
        ("axes", C.c_int64 * MAX_RANK),
        ("n_axes", C.c_int32),
        ("axis", C.c_int64),
        ("momentum", C.c_double),
        ("seed", C.c_uint64),
        ("math", C.c_int32),
        ("reserved", C.c_int32 * 7),
    ]


class Param(C.Structure):
    _fields_ = [("w", C.c_void_p), ("g", C.c_void_p), ("s0", C.c_void_p), ("s1", C.c_void_p), ("s2", C.c_void_p),
                ("n", C.c_int64)]


class Jitter(C.Structure):
    _fields_ = [("x_off", C.c_int32), ("y_off", C.c_int32), ("flip_x", C.c_int32), ("flip_y", C.c_int32)]


# every symbol include/dopt_b200.h declares; tests/test_abi.py checks the library exports each of them
SYMBOLS = [
    "dopt_b200_init", "dopt_b200_last_error", "dopt_b200_version", "dopt_b200_device_info",
    "dopt_b200_set_default_math", "dopt_b200_launch_count", "dopt_b200_tc_profile", "dopt_b200_list_operations", "dopt_b200_has_operation",
    "dopt_b200_kernel_create", "dopt_b200_kernel_execute", "dopt_b200_kernel_destroy",
    "dopt_b200_sgd_update", "dopt_b200_adam_update",
    "dopt_b200_image_transform_u8", "dopt_b200_image_transform_f32", "dopt_b200_one_hot_u8", "dopt_b200_jitter_sample",
    "dopt_b200_plan_create", "dopt_b200_plan_add_node", "dopt_b200_plan_set_outputs", "dopt_b200_plan_finalize",
    "dopt_b200_plan_execute", "dopt_b200_plan_stats", "dopt_b200_plan_profile", "dopt_b200_plan_replay_class", "dopt_b200_plan_destroy",
    "dopt_b200_comm_unique_id", "dopt_b200_comm_init", "dopt_b200_comm_world_size", "dopt_b200_comm_rank",
    "dopt_b200_allreduce", "dopt_b200_comm_check", "dopt_b200_comm_destroy", "dopt_b200_comm_set_symmetric",
]


def load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libdopt_b200.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.dopt_b200_init.restype = C.c_int
    lib.dopt_b200_last_error.restype = C.c_char_p
    lib.dopt_b200_version.restype = C.c_char_p
    lib.dopt_b200_device_info.argtypes = [C.POINTER(C.c_int)] * 3 + [C.POINTER(C.c_size_t)]
    lib.dopt_b200_set_default_math.argtypes = [C.c_int]
    lib.dopt_b200_set_default_math.restype = None
    lib.dopt_b200_launch_count.restype = C.c_uint64
    lib.dopt_b200_tc_profile.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.dopt_b200_list_operations.restype = C.POINTER(C.c_char)
    lib.dopt_b200_has_operation.argtypes = [C.c_char_p]
    lib.dopt_b200_kernel_create.argtypes = [C.POINTER(Op), C.POINTER(vp)]
    lib.dopt_b200_kernel_execute.argtypes = [vp, C.POINTER(vp), C.c_int, vp, vp]
    lib.dopt_b200_kernel_destroy.argtypes = [vp]
    fp = vp
    lib.dopt_b200_sgd_update.argtypes = [C.POINTER(Param), C.c_int, fp, fp, C.c_int, C.c_float, vp]
    lib.dopt_b200_adam_update.argtypes = [C.POINTER(Param), C.c_int, fp, fp, fp, fp, fp, fp, C.c_int, C.c_float, vp]
    img = [vp, vp, i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]
    lib.dopt_b200_image_transform_u8.argtypes = img
    lib.dopt_b200_image_transform_f32.argtypes = img
    lib.dopt_b200_one_hot_u8.argtypes = [vp, vp, i64, C.c_int, vp]
    lib.dopt_b200_jitter_sample.argtypes = [vp, i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint64, vp]
    lib.dopt_b200_plan_create.argtypes = [C.POINTER(vp)]
    lib.dopt_b200_plan_add_node.argtypes = [vp, C.POINTER(Op), C.POINTER(i32), C.c_int, vp]
    lib.dopt_b200_plan_set_outputs.argtypes = [vp, C.POINTER(i32), C.c_int]
    lib.dopt_b200_plan_finalize.argtypes = [vp, C.c_int]
    lib.dopt_b200_plan_execute.argtypes = [vp, C.POINTER(i32), C.POINTER(vp), C.POINTER(i32), C.c_int, C.POINTER(vp),
                                           C.c_int, vp]
    lib.dopt_b200_plan_stats.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]
    lib.dopt_b200_plan_profile.argtypes = [vp, C.c_int, C.c_char_p, C.c_size_t]
    lib.dopt_b200_plan_replay_class.argtypes = [vp, C.c_char_p, C.c_int, C.POINTER(C.c_double), C.POINTER(i64), vp]
    lib.dopt_b200_plan_destroy.argtypes = [vp]
    lib.dopt_b200_comm_unique_id.argtypes = [vp]
    lib.dopt_b200_comm_init.argtypes = [C.c_int, C.c_int, vp]
    lib.dopt_b200_allreduce.argtypes = [vp, i64, C.c_float, vp]
    lib.dopt_b200_comm_set_symmetric.argtypes = [vp, vp, C.c_size_t, C.POINTER(vp), C.c_int, C.c_size_t]
    return lib
