#!/usr/bin/env python
"""Generates tests/golden/cudnn_golden.npz ON THE GPU BOX: small seeded inputs pushed through the reference CUDA backend's
own cuDNN / cuBLAS call sequences (oracle/cudnn_replay.cpp replays cuda/source/dopt/cuda/nnet/cudnn7.d, math.d, basic.d),
inputs and outputs stored side by side.  The CPU test suite (tests/test_oracle_vs_cudnn_golden.py) pins the numpy oracle
against these vectors without needing a GPU.

    gpurun -- 'python tools/make_cudnn_golden.py gpurun_out/cudnn_golden.npz'   then copy the file to tests/golden/
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cudnn_replay as Q  # noqa: E402

F = np.float32
out = {}
rng = np.random.RandomState(20261017)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    torch.cuda.synchronize()
    return t.cpu().numpy()


def put(prefix, **arrays):
    for k, v in arrays.items():
        out["%s/%s" % (prefix, k)] = np.asarray(v)


v = Q.versions()
put("meta", cudnn=np.int64(v["cudnn"]), cudart=np.int64(v["cudart"]))

# ---- convolution family (CUDNN_CONVOLUTION = flipped filters; CUDNN_FMA_MATH) ----
for i, (N, C, H, W, K, R, S, pad, stride) in enumerate([(2, 3, 8, 8, 4, 3, 3, 1, 1), (3, 8, 9, 7, 6, 3, 2, 1, 2),
                                                        (2, 1, 12, 12, 4, 5, 5, 0, 1), (2, 6, 8, 8, 5, 1, 1, 0, 2)]):
    x = rng.randn(N, C, H, W).astype(F)
    w = (rng.randn(K, C, R, S) * np.sqrt(2.0 / (C * R * S))).astype(F)
    p, s = [pad, pad], [stride, stride]
    y = Q.convolution(dev(x), dev(w), p, s, math=1)
    dy = rng.randn(*y.shape).astype(F)
    dx = Q.convolution_features_grad(dev(dy), dev(w), x.shape, p, s, math=1)
    dw = Q.convolution_filters_grad(dev(dy), dev(x), w.shape, p, s, math=1)
    put("conv%d" % i, x=x, w=w, dy=dy, padding=np.array(p), stride=np.array(s), y=host(y), dx=host(dx), dw=host(dw))

# ---- max pooling, with tied windows (zeros after a relu) ----
for i, (shape, dims) in enumerate([((2, 3, 6, 6), [2, 2]), ((3, 5, 9, 7), [2, 3])]):
    x = np.maximum(rng.randn(*shape), 0).astype(F)          # about half the entries are exact zeros
    x[0, 0, :2, :2] = 1.25                                   # a fully tied non-zero window
    y = Q.maxpool(dev(x), dims)
    dy = rng.randn(*y.shape).astype(F)
    dx = Q.maxpool_grad(dev(dy), y, dev(x), dims)
    put("pool%d" % i, x=x, dims=np.array(dims), y=host(y), dy=dy, dx=host(dx))

# ---- softmax / relu / bias ----
for i, shape in enumerate([(4, 7), (2, 5, 3, 2)]):
    x = (rng.randn(*shape) * 3).astype(F)
    y = Q.softmax(dev(x))
    dy = rng.randn(*shape).astype(F)
    dx = Q.softmax_grad(dev(dy), y)
    put("softmax%d" % i, x=x, y=host(y), dy=dy, dx=host(dx))
x = rng.randn(3, 4, 5).astype(F)
x.reshape(-1)[::5] = 0.0
y = Q.relu(dev(x))
dy = rng.randn(3, 4, 5).astype(F)
put("relu", x=x, y=host(y), dy=dy, dx=host(Q.relu_grad(dev(dy), y, dev(x))))
x = rng.randn(3, 4, 5, 2).astype(F)
b = rng.randn(4).astype(F)
put("bias", x=x, b=b, y=host(Q.add_bias(dev(x), dev(b))), db=host(Q.add_bias_grad(dev(x))),
    db_second_execution=host(Q.add_bias_grad(dev(x), executions=2)))

# ---- batch norm: spatial (rank 4) and rank 2 ----
for i, shape in enumerate([(6, 5, 7, 3), (10, 8)]):
    C = shape[1]
    x = (rng.randn(*shape) * 1.3 - 0.4).astype(F)
    g = (rng.rand(C) + 0.5).astype(F)
    b = rng.randn(C).astype(F)
    rm = rng.randn(C).astype(F)
    rv = (rng.rand(C) + 0.5).astype(F)
    dy = rng.randn(*shape).astype(F)
    put("bn%d" % i, x=x, scale=g, bias=b, mean=rm, var=rv, dy=dy, momentum=np.float64(0.9),
        train=host(Q.batch_norm_train(dev(x), dev(g), dev(b), dev(rm), dev(rv), 0.9)),
        grad=host(Q.batch_norm_grad(dev(dy), dev(x), dev(g))),
        inference=host(Q.batch_norm_inference(dev(x), dev(g), dev(b), dev(rm), dev(rv))))

# ---- cuBLAS ----
a, b = rng.randn(5, 7).astype(F), rng.randn(7, 3).astype(F)
put("matmul", a=a, b=b, c=host(Q.matmul(dev(a), dev(b))))
a = rng.randn(4, 6).astype(F)
put("transpose", a=a, c=host(Q.transpose(dev(a))))

path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "cudnn_golden.npz")
np.savez_compressed(path, **out)
print("wrote %s: %d arrays, cuDNN %d" % (path, len(out), v["cudnn"]))
