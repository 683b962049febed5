"""Per-kernel micro-benchmarks on the GPU box: CUDA-event timing of single ops at BASELINE shapes, reported against the
measured roofline (MEASURED_PEAKS.json).  Usage: python tools/bench_ops.py [conv|bw|all]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dopt_b200 as db  # noqa: E402

PEAKS = {"hbm_gbs": 6555.5, "bf16_tflops": 1676.7}
try:
    PEAKS.update(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))))
except Exception:
    pass


def tc_profile(enable):
    """(us, launches) of tc_kernel since the last tc_profile(True)"""
    import ctypes as C
    us, n = C.c_double(0), C.c_int64(0)
    db.lib.dopt_b200_tc_profile(int(enable), C.byref(us), C.byref(n))
    return us.value, n.value


def time_kernel(k, ins, out, iters=20, warm=3, flush=None):
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(warm):
        k.execute(ins, out, s)
    torch.cuda.synchronize()
    ts = []
    tc_profile(True)
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        k.execute(ins, out, s)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def conv_bench():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    cases = [(128, 160, 32, 160, 3, 1, 1), (128, 320, 16, 320, 3, 1, 1), (128, 640, 8, 640, 3, 1, 1),
             (128, 160, 32, 320, 3, 1, 2), (128, 16, 32, 160, 3, 1, 1), (128, 160, 32, 320, 1, 0, 2)]
    if os.environ.get("BENCH_OPS_FIRST"):
        cases = cases[:int(os.environ["BENCH_OPS_FIRST"])]
    for (N, C, H, K, R, pad, st) in cases:
        P = (H + 2 * pad - R) // st + 1
        x = torch.randn(N, C, H, H, device="cuda")
        w = torch.randn(K, C, R, R, device="cuda") * 0.05
        y = torch.empty(N, K, P, P, device="cuda")
        dy = torch.randn(N, K, P, P, device="cuda")
        dx = torch.empty_like(x)
        dw = torch.empty_like(w)
        flops = 2.0 * N * K * C * R * R * P * P
        at = dict(padding=[pad, pad], stride=[st, st])
        for name, ins, out, shapes in [
            ("convolution", [x, w], y, ([x.shape, w.shape], y.shape)),
            ("convolutionFeaturesGrad", [dy, w], dx, ([dy.shape, w.shape], x.shape)),
            ("convolutionFiltersGrad", [dy, x], dw, ([dy.shape, x.shape], w.shape)),
        ]:
            op = db.make_op(name, [tuple(s) for s in shapes[0]], tuple(shapes[1]), **at)
            k = db.CUDAKernel(op)
            ms = time_kernel(k, ins, out, flush=flush)
            us, n = tc_profile(False)
            kms = us / 1e3 / 20 if n else float("nan")   # tc_kernel alone (all its launches of one op execution)
            print("%-24s N%d C%d H%d K%d R%d s%d: op %6.3f ms | tc_kernel %6.3f ms %6.1f TFLOP/s (%4.1f%% of bf16 peak)"
                  % (name, N, C, H, K, R, st, ms, kms, flops / kms / 1e9, 100 * flops / kms / 1e9 / PEAKS["bf16_tflops"]))
            k.close()


def bw_bench():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    V = 128 * 160 * 32 * 32
    x = torch.randn(128, 160, 32, 32, device="cuda")
    y = torch.empty_like(x)
    z = torch.empty_like(x)

    def rep(name, k, ins, out, nbytes):
        ms = time_kernel(k, ins, out, flush=flush)
        gbs = nbytes / ms / 1e6
        print("%-22s %8.3f ms  %7.1f GB/s (%.1f%% of measured HBM peak)" % (name, ms, gbs, 100 * gbs / PEAKS["hbm_gbs"]))

    sh = tuple(x.shape)
    rep("add", db.CUDAKernel(db.make_op("add", [sh, sh], sh)), [x, y], z, 3 * V * 4)
    rep("relu", db.CUDAKernel(db.make_op("relu", [sh], sh)), [x], z, 2 * V * 4)
    rep("reluGrad", db.CUDAKernel(db.make_op("reluGrad", [sh, sh, sh], sh)), [x, y, y], z, 3 * V * 4)
    C = 160
    g = torch.ones(1, C, 1, 1, device="cuda")
    b = torch.zeros(C, device="cuda")
    packed = torch.empty(V + 2 * C, device="cuda")
    rep("batchNormTrain", db.CUDAKernel(db.make_op("batchNormTrain", [sh, (1, C, 1, 1), (C,), (C,), (C,)], (V + 2 * C,),
                                                   momentum=0.9)), [x, g, b, b, b + 1], packed, 2 * V * 4)
    packed2 = torch.empty(2 * V + C, device="cuda")
    rep("batchNormGrad", db.CUDAKernel(db.make_op("batchNormGrad", [sh, sh, (1, C, 1, 1)], (2 * V + C,))), [y, x, g],
        packed2, 3 * V * 4)
    w = torch.randn(640, 640, 3, 3, device="cuda")
    out1 = torch.empty(1, device="cuda")
    rep("sum(3.7M)", db.CUDAKernel(db.make_op("sum", [tuple(w.shape)], (), axes=[0, 1, 2, 3])), [w], out1, w.numel() * 4)
    a = torch.randn(V, 1, device="cuda")
    s1 = torch.randn(1, 1, device="cuda")
    rep("matmul [V,1]x[1,1]", db.CUDAKernel(db.make_op("matmul", [(V, 1), (1, 1)], (V, 1))), [a, s1], z, 2 * V * 4)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    db.check(db.lib.dopt_b200_init())
    if what in ("conv", "all"):
        conv_bench()
    if what in ("bw", "all"):
        bw_bench()
