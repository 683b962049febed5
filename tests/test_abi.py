"""The drop-in boundary: libdopt_b200.so loads, exports every symbol include/dopt_b200.h declares, registers a kernel for
every op type dopt's CUDA backend registers, and fails loudly (never falls back) when no GPU is present."""
import ctypes
import os
import re

import pytest

import dopt_b200 as db
from dopt_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# cuda/source/dopt/cuda/math.d:79-99, basic.d:13-22, nnet/cudnn7.d:14-34, random.d (registerCUDAKernel calls)
REFERENCE_CUDA_OPS = [
    "add", "sub", "mul", "div", "lt", "lte", "gt", "gte", "eq", "neq", "max", "min", "pow",
    "neg", "abs", "sgn", "exp", "log", "sqrt", "matmul",
    "slice", "pad", "repeat", "transpose",
    "convolution", "convolutionFeaturesGrad", "convolutionFiltersGrad", "maxpool", "maxpoolGrad", "softmax",
    "softmaxGrad", "relu", "reluGrad", "addBias", "addBiasGrad", "batchNormTrain", "batchNormGrad",
    "batchNormInference", "uniform",
]
# ops the reference leaves to its D2H -> CPU -> H2D fallback (cuda/source/dopt/cuda/package.d:81-119,284)
FALLBACK_OPS_NOW_ON_GPU = ["sum", "maxElement", "argmin"]


def test_header_symbols_are_exported():
    header = open(os.path.join(ROOT, "include", "dopt_b200.h")).read()
    declared = set(re.findall(r"\b(dopt_b200_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found in the header"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert getattr(raw, name) is not None, name


def test_d_glue_declares_only_what_the_header_declares():
    """dopt_b200/d/dopt/b200/package.d cannot be compiled here (no D toolchain): at least its extern(C) block must stay in step
    with the header -- every function it names exists there with the same number of parameters, and every entry point a dopt
    host needs (kernels, plans, communicator, input pipeline, math mode) is declared in it."""
    def functions(text):
        text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)       # comments mention entry points too
        text = re.sub(r"//[^\n]*", " ", text)
        out = {}
        for m in re.finditer(r"\b(dopt_b200_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, re.S):
            args = m.group(2).strip()
            out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
        return out
    header = functions(open(os.path.join(ROOT, "include", "dopt_b200.h")).read())
    glue_src = open(os.path.join(ROOT, "dopt_b200", "d", "dopt", "b200", "package.d")).read()
    glue = functions(glue_src[glue_src.index("extern(C)"):glue_src.index("enum DOPT_B200_PLAN_FUSE")])
    assert len(glue) >= 30
    for name, n_args in glue.items():
        assert name in header, name
        assert header[name] == n_args, (name, header[name], n_args)
    missing = set(header) - set(glue) - {"dopt_b200_launch_count", "dopt_b200_tc_profile"}    # bench.py's counters
    assert not missing, missing


def test_every_reference_op_has_a_kernel():
    ops = set(db.list_operations())
    for name in REFERENCE_CUDA_OPS + FALLBACK_OPS_NOW_ON_GPU:
        assert name in ops, name
        assert db.lib.dopt_b200_has_operation(name.encode()) == 1
    assert db.lib.dopt_b200_has_operation(b"noSuchOp") == 0


def test_struct_layout_matches_header():
    # sizes the C compiler gives the PODs (computed from the header's field list)
    assert ctypes.sizeof(_lib.Tensor) == 4 + 4 + 8 * 8
    assert ctypes.sizeof(_lib.Param) == 6 * 8
    assert ctypes.sizeof(_lib.Op) % 8 == 0


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    op = db.make_op("add", [(4,), (4,)], (4,))
    with pytest.raises(db.DoptError) as e:
        db.CUDAKernel(op)
    assert "no CPU fallback" in str(e.value)
    assert db.lib.dopt_b200_init() != 0


def test_unknown_op_is_an_error_not_a_fallback():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a device to get past the device check")
    op = db.make_op("sinh", [(4,)], (4,))
    with pytest.raises(db.DoptError) as e:
        db.CUDAKernel(op)
    assert "Could not construct a CUDA kernel" in str(e.value)


def test_ctypes_layout_equals_c_layout(tmp_path):
    """Compile a C probe against the real header and compare sizeof/offsetof with the ctypes mirror."""
    import subprocess
    src = tmp_path / "probe.c"
    fields = ["op_type", "n_inputs", "inputs", "output", "padding", "stride", "pool_dims", "start", "stop", "before",
              "after", "repetitions", "order", "axes", "n_axes", "axis", "momentum", "seed", "math"]
    body = "".join('printf("%s %%zu\\n", offsetof(dopt_b200_op, %s));\n' % (f, f) for f in fields)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "dopt_b200.h"\nint main(){\n'
                   'printf("sizeof_op %zu\\n", sizeof(dopt_b200_op));\n'
                   'printf("sizeof_tensor %zu\\n", sizeof(dopt_b200_tensor));\n'
                   'printf("sizeof_param %zu\\n", sizeof(dopt_b200_param));\n' + body + 'return 0;}\n')
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = dict(line.split() for line in subprocess.check_output([str(exe)]).decode().splitlines())
    assert int(out["sizeof_op"]) == ctypes.sizeof(_lib.Op)
    assert int(out["sizeof_tensor"]) == ctypes.sizeof(_lib.Tensor)
    assert int(out["sizeof_param"]) == ctypes.sizeof(_lib.Param)
    for f in fields:
        assert int(out[f]) == getattr(_lib.Op, f).offset, f


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under dopt_b200/ (Python, C++, CUDA, D) may import, include, link or load it."""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "dopt_b200")):
        if os.sep + "build" in base or os.sep + "lib" in base or "__pycache__" in base:
            continue
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".d", "Makefile")):
                continue
            text = open(os.path.join(base, f), errors="replace").read()
            # code references only (comments may cite test files): imports, includes, link flags, dlopen targets
            for needle in ("from oracle", "import oracle", "#include <cudnn", "#include \"cudnn", "oracle/cudnn_replay",
                           "libcudnn_replay", "-lcudnn", "\"libcudnn"):
                if needle in text:
                    bad.append((os.path.join(base, f), needle))
    assert not bad, bad
