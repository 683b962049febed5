"""Training throughput of the BASELINE.json configurations that are parity cases rather than bench lines (configs[2] VGG19 + batch
norm at 100 x 3 x 32 x 32 with SGD; configs[4] SINS Wide ResNet-16-8 at 50 x 3 x 96 x 96, strides [2,2,2], AMSGrad), end to end through
`updater(args)` with host arrays, default (production) plan mode.  One JSON line per configuration; not a bench contract line."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from dopt_b200 import host as H   # noqa: E402

F = np.float32


def run(name, build, kind, hyper, batch, steps=20, warmup=5):
    H.reset()
    H.seed(1)
    x, y, preds = build()
    net = H.Network([x], [preds])
    loss = H.cross_entropy(preds.train_output, y) + net.param_loss
    upd = H.Updater(kind, [loss, preds.train_output], network=net, hyper=hyper())
    rng = np.random.RandomState(0)
    data = [((rng.rand(*x.shape) * 2 - 1).astype(F), np.eye(y.shape[1], dtype=F)[rng.randint(0, y.shape[1], batch)]) for _ in range(4)]
    losses = []
    for s in range(warmup):
        losses.append(float(upd.step({x: data[s % 4][0], y: data[s % 4][1]})[0]))
    t0 = time.perf_counter()
    for s in range(steps):
        losses.append(float(upd.step({x: data[s % 4][0], y: data[s % 4][1]})[0]))
    dt = (time.perf_counter() - t0) / steps
    st = upd.stats()
    print(json.dumps({"config": name, "images_per_s": batch / dt, "ms_per_step": dt * 1e3, "batch": batch,
                      "params": int(sum(p.volume for p in net.params)), "launches_per_step": st["launches"],
                      "plan_device_bytes": st["device_bytes"], "loss_first": losses[0], "loss_last": losses[-1],
                      "how": "updater(args) with pageable host arrays: H2D, one CUDA-graph step, D2H of loss and predictions"}))


def main():
    assert H.init(), H.init_error()

    def vgg():
        x, y = H.float32((100, 3, 32, 32)), H.float32((100, 10))
        return x, y, H.vgg19(x, dense_sizes=(512, 512), batchnorm=True).dense(10).softmax()

    def sins():
        x, y = H.float32((50, 3, 96, 96)), H.float32((50, 10))
        return x, y, H.wide_resnet(x, 16, 8, stride=(2, 2, 2)).dense(10).softmax()
    run("configs[2] VGG19+BN 100x3x32x32 SGD", vgg, H.SGD, lambda: [H.float32((), [0.01]), H.float32((), [0.9])], 100)
    run("configs[4] SINS WRN-16-8 50x3x96x96 AMSGrad", sins, H.AMSGRAD, lambda: [H.float32((), [1e-4]), None, None, None], 50)


if __name__ == "__main__":
    main()
