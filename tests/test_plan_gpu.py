"""Whole-graph parity on the GPU: graphs built by the dopt mirror, executed by (a) the reference-style node-by-node
CUDAPlan over the per-op C ABI and (b) the lowered / fused / CUDA-graph B200Plan, against the graph-level CPU oracle.
Includes the reference's own end-to-end unit tests and multi-step loss curves for the BASELINE model families at sizes the
oracle finishes in seconds.

Tolerances: fp32 math -> rtol 1e-4 per step on losses (summation order); bf16 tensor-core math -> losses within 2e-2
relative, parameters within 3e-2 of the largest update-scaled magnitude after a few steps.  Where `update_tol` is given
the check is per parameter tensor on the UPDATE it received, ||(p_gpu - p_0) - (p_oracle - p_0)||_2 / ||p_oracle - p_0||_2,
with one bound per parameter class (a wrong gradient on any tensor, however small its values, shows up as O(1))."""
import numpy as np
import pytest

import dopt_b200 as db
from dopt_b200 import host as H
from oracle import graph_eval as G

pytestmark = pytest.mark.gpu
F = np.float32
FUSE, GRAPH, INTERIOR = db._lib.PLAN_FUSE, db._lib.PLAN_CUDA_GRAPH, db._lib.PLAN_BF16_INTERIOR
MODES = [("cudaplan", 1, 0), ("plain", 0, 0), ("fused", 0, FUSE), ("fused+graph", 0, FUSE | GRAPH)]


@pytest.fixture(autouse=True)
def _fresh():
    assert H.init(), H.init_error()
    H.reset()
    H.set_math(db.MATH_FP32)
    H.set_plan_flags(FUSE | GRAPH)
    yield
    H.reset()
    H.set_math(db.MATH_DEFAULT)


def run_modes(outputs, args=None, reps=1):
    res = {}
    for name, kind, flags in MODES:
        H.set_plan_flags(flags)
        p = H.Plan(outputs, kind=kind)
        for _ in range(reps):
            out = p.execute(args)
        res[name] = out
    return res


def test_reference_cuda_backend_unittest():
    # cuda/source/dopt/cuda/package.d:533-542
    a, b, c = H.float32((), [3.0]), H.float32((), [4.0]), H.float32((), [-1.0])
    y = a * b + c
    for name, out in run_modes([y], reps=3).items():
        assert out[0] == F(11.0), name


def test_core_kats_on_gpu():
    m = H.float32((2, 2), [0, 1, 2, 5])
    outs = [H.matmul(H.float32((2, 1), [1, 2]), H.float32((1, 2), [3, 4])), H.sum_(m), H.sum_(m, [0]), H.sum_(m, [1]),
            H.argmin(H.float32((2, 3), [5, 1, 3, 6, 7, 2]), 1), H.max_element(m, [0]),
            H.repeat(H.float32((2,), [1, 2]), 3), H.transpose(m, [1, 0]),
            H.convolution(H.float32((1, 1, 3, 5), [1, 1, 1, 0, 0] * 3), H.float32((1, 1, 1, 2), [-1, 1])),
            H.softmax(H.float32((1, 5), [1, 2, 3, 1, 2]))]
    want = [[3, 4, 6, 8], [8], [2, 6], [1, 7], [1, 2], [2, 5], [1, 2, 1, 2, 1, 2], [0, 2, 1, 5], [0, 0, 1, 0] * 3, None]
    for name, out in run_modes(outs, reps=3).items():
        for o, w in zip(out, want):
            if w is not None:
                assert o.ravel().tolist() == w, name
        np.testing.assert_allclose(out[-1].ravel(), [0.0674508, 0.18335, 0.498398, 0.0674508, 0.18335], atol=1e-6)


def test_autodiff_kats_on_gpu():
    rng = np.random.RandomState(0)
    xv = rng.randn(3, 4).astype(F)
    x = H.float32((3, 4), xv)
    g = H.grad(H.sum_(x * x), [x])[0]
    a = H.float32((4, 4), [1, 2, 3, 4] * 4)
    b = H.float32((4, 4), [5, 6, 7, 8] * 4)
    ga = H.grad(H.slice_(a * b, [1, 1], [2, 2]), [a])[0]
    for name, out in run_modes([g, ga], reps=2).items():
        np.testing.assert_array_equal(out[0], 2 * xv)      # (1*x) + (1*x): exact
        assert out[1].ravel().tolist() == [0, 0, 0, 0, 0, 6, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0], name


def test_lowering_removes_bn_scaffolding_and_broadcasts():
    H.seed(5)
    x = H.float32((8, 16, 8, 8))
    y = H.float32((8, 10))
    out = H.data_source(x).conv2d(16, (3, 3), padding=(1, 1), use_bias=False, weight_decay=1e-4).batch_norm().relu() \
        .dense(10).softmax()
    net = H.Network([x], [out])
    loss = H.cross_entropy(out.train_output, y) + net.param_loss
    stats = {}
    for name, kind, flags in MODES[1:]:
        H.set_plan_flags(flags)
        upd = H.Updater(H.SGD, [loss], network=net, hyper=[H.float32((), [0.1]), H.float32((), [0.9])])
        rng = np.random.RandomState(1)
        upd.step({x: rng.randn(8, 16, 8, 8).astype(F), y: np.eye(10, dtype=F)[rng.randint(0, 10, 8)]})
        stats[name] = upd.stats()
    assert stats["fused"]["lowered_nodes"] < stats["plain"]["lowered_nodes"]
    assert stats["fused"]["launches"] < stats["plain"]["launches"]


def _param_class(p):
    if len(p.shape) == 4 and p.shape[0] != 1:
        return "conv"
    if len(p.shape) == 2:
        return "dense"
    return "bn_or_bias"


def _update_errors(params, initial, oracle):
    """per parameter class: worst ||(p_gpu - p_0) - (p_oracle - p_0)||_2 / ||p_oracle - p_0||_2 over the tensors of the class"""
    worst = {}
    for p, p0 in zip(params, initial):
        gv, wv = p.get(), oracle.value_of(p)
        du = np.linalg.norm((wv - p0).astype(np.float64))
        if du <= 1e-12:       # a tensor nothing updated (zero gradient): nothing to compare
            continue
        err = float(np.linalg.norm((gv - wv).astype(np.float64)) / du)
        cls = _param_class(p)
        if err > worst.get(cls, (0.0, None))[0]:
            worst[cls] = (err, tuple(p.shape))
    return worst


def _train_compare(build, steps, math, loss_rtol, param_tol, flags=FUSE | GRAPH, kind=H.SGD, hyper=None, update_tol=None,
                   report=None, first_step_tol=None):
    """build() -> (loss op, extra outputs, network, feed function).  Steps the GPU updater and the oracle side by side."""
    H.set_math(math)
    H.set_plan_flags(flags)
    loss, extra, net, feed = build()
    hyper = hyper if hyper is not None else [H.float32((), [0.05]), H.float32((), [0.9])]
    upd = H.Updater(kind, [loss] + extra, network=net, hyper=hyper)
    oracle = G.UpdaterOracle(upd)
    initial = [p.get().copy() for p in net.params]
    losses = []
    for s in range(steps):
        args = feed(s)
        got = upd.step(args)
        want = oracle.step(args)
        losses.append((float(got[0]), float(want[0])))
        assert abs(got[0] - want[0]) <= loss_rtol * max(1.0, abs(float(want[0]))), (s, losses)
        for g, w in zip(got[1:], want[1:]):
            assert np.abs(g - w).max() <= max(loss_rtol * 10, 1e-5) * max(1.0, float(np.abs(w).max())), s
        if s == 0 and first_step_tol is not None:
            # after one step from identical parameters the update IS the gradient (times the learning rate): a per-tensor
            # gradient parity check, untouched by the divergence of two trajectories
            first = _update_errors(net.params, initial, oracle)
            if report is not None:
                report["first_step"] = first
            for cls, (err, shape) in first.items():
                assert err <= first_step_tol[cls], ("first step", cls, shape, err, first_step_tol[cls])
    if param_tol is not None:
        for p in net.params:
            gv, wv = p.get(), oracle.value_of(p)
            scale = max(float(np.abs(wv).max()), 1e-3)
            assert float(np.abs(gv - wv).max()) <= param_tol * scale, (p.shape, float(np.abs(gv - wv).max()), scale)
    if update_tol is not None:
        worst = _update_errors(net.params, initial, oracle)
        if report is not None:
            report["all_steps"] = worst
        for cls, (err, shape) in worst.items():
            assert err <= update_tol[cls], ("all steps", cls, shape, err, update_tol[cls])
    return losses


def _small_convnet(batch=8, bias=True, bn=True, pool=True):
    def build():
        H.seed(11)
        x = H.float32((batch, 16, 8, 8))
        y = H.float32((batch, 10))
        l = H.data_source(x).conv2d(32, (3, 3), padding=(1, 1), weight_decay=1e-3, use_bias=bias)
        if bn:
            l = l.batch_norm()
        l = l.relu()
        if pool:
            l = l.max_pool((2, 2))
        l = l.conv2d(32, (3, 3), padding=(1, 1), stride=(2, 2), weight_decay=1e-3, use_bias=bias).relu().dense(10).softmax()
        net = H.Network([x], [l])
        loss = H.cross_entropy(l.train_output, y) + net.param_loss
        rng = np.random.RandomState(3)
        data = [(rng.randn(batch, 16, 8, 8).astype(F), np.eye(10, dtype=F)[rng.randint(0, 10, batch)]) for _ in range(8)]
        return loss, [l.train_output], net, lambda s: {x: data[s % 8][0], y: data[s % 8][1]}
    return build


@pytest.mark.parametrize("mode", MODES[1:], ids=[m[0] for m in MODES[1:]])
def test_small_convnet_sgd_fp32_all_plan_modes(mode):
    _train_compare(_small_convnet(), 5, db.MATH_FP32, 2e-4, 2e-3, flags=mode[2])


def test_small_convnet_sgd_bf16():
    _train_compare(_small_convnet(batch=16), 5, db.MATH_BF16, 2e-2, 5e-2)


@pytest.mark.parametrize("kind", [H.ADAM, H.AMSGRAD, H.SGD_NESTEROV])
def test_small_convnet_other_optimisers(kind):
    hyper = [H.float32((), [1e-3]), None, None, None] if kind in (H.ADAM, H.AMSGRAD) else None
    _train_compare(_small_convnet(bn=False), 4, db.MATH_FP32, 2e-4, 5e-3, kind=kind, hyper=hyper)


def test_mnist_logistic_regression_adam():
    # examples/mnistlogit.d:34-51: softmax regression written by hand ([100,784] x [784,10]); BASELINE configs[0]
    def build():
        rng = np.random.RandomState(4)
        x, y = H.float32((100, 784)), H.float32((100, 10))
        W = H.float32((784, 10), (rng.randn(784, 10) * 0.01).astype(F))
        b = H.float32((10,))
        logits = H.matmul(x, W) + H.repeat(b, 100)
        e = H.unary("exp", logits)
        denom = H.repeat(H.reshape(H.sum_(e, [1]), [100, 1]), [1, 10])
        p = e / denom
        loss = H.sum_(y * H.unary("log", p + 1e-6)) * (-1.0 / 100)
        data = [(rng.rand(100, 784).astype(F), np.eye(10, dtype=F)[rng.randint(0, 10, 100)]) for _ in range(4)]

        class Net(object):
            params = [W, b]
            h = -1
        return loss, [], Net, lambda s: {x: data[s % 4][0], y: data[s % 4][1]}
    H.set_math(db.MATH_FP32)
    loss, extra, net, feed = build()
    upd = H.Updater(H.ADAM, [loss], wrt=net.params, hyper=[H.float32((), [1e-3]), None, None, None])
    oracle = G.UpdaterOracle(upd)
    for s in range(6):
        got, want = upd.step(feed(s)), oracle.step(feed(s))
        assert abs(got[0] - want[0]) < 2e-4 * abs(float(want[0]))
    for p in net.params:
        np.testing.assert_allclose(p.get(), oracle.value_of(p), rtol=0, atol=2e-5)


def test_mnist_cnn_adam():
    # examples/mnist.d:35-58: conv5x5(32)-relu-pool-conv5x5(32)-relu-pool-dense(10)-softmax, Adam 1e-3 (batch reduced)
    def build():
        H.seed(12)
        x, y = H.float32((10, 1, 28, 28)), H.float32((10, 10))
        l = H.data_source(x).conv2d(32, (5, 5)).relu().max_pool((2, 2)).conv2d(32, (5, 5)).relu().max_pool((2, 2)) \
            .dense(10).softmax()
        net = H.Network([x], [l])
        loss = H.cross_entropy(l.train_output, y) + net.param_loss
        rng = np.random.RandomState(5)
        data = [(rng.rand(10, 1, 28, 28).astype(F), np.eye(10, dtype=F)[rng.randint(0, 10, 10)]) for _ in range(3)]
        return loss, [l.train_output], net, lambda s: {x: data[s % 3][0], y: data[s % 3][1]}
    _train_compare(build, 3, db.MATH_FP32, 2e-4, 5e-3, kind=H.ADAM, hyper=[H.float32((), [1e-3]), None, None, None])


def _wrn(depth, width, batch, hw, classes, strides=(1, 2, 2)):
    def build():
        H.seed(13)
        x, y = H.float32((batch, 3, hw, hw)), H.float32((batch, classes))
        preds = H.wide_resnet(x, depth, width, stride=strides).dense(classes).softmax()
        net = H.Network([x], [preds])
        loss = H.cross_entropy(preds.train_output, y) + net.param_loss
        rng = np.random.RandomState(6)
        data = [((rng.rand(batch, 3, hw, hw) * 2 - 1).astype(F), np.eye(classes, dtype=F)[rng.randint(0, classes, batch)])
                for _ in range(4)]
        return loss, [preds.train_output], net, lambda s: {x: data[s % 4][0], y: data[s % 4][1]}
    return build


def test_wrn_16_2_sgd_fp32():
    # the cifar100.d recipe (SGD lr 0.1 momentum 0.9, wd 1e-4) on a narrow WRN at 16x16 so the oracle stays fast
    _train_compare(_wrn(16, 2, 4, 16, 10), 3, db.MATH_FP32, 5e-4, 1e-2,
                   hyper=[H.float32((), [0.1]), H.float32((), [0.9])])


# per-class bounds on the relative error of the update a parameter tensor received (see the module docstring).
# FIRST_*: after the first step (= gradient parity per tensor).  UPD_*: after all steps of the run, where the two trajectories
# have started to drift apart (batch 4..32 with lr 0.05..0.1 amplifies rounding differences step by step).
# Measured on a B200 (profiles/r02_parity.md), WRN-28-10 at batch 4 -- the hardest case: 28 layers deep, batch-norm statistics
# over as few as 256 samples, whose backward pass subtracts two nearly equal sums -- first step / after three steps:
#   fp32                      conv 1.6e-3 / 1.8e-2   bn 2.3e-3 / 2.1e-2   dense 5e-7 / 8e-5
#   bf16 operands             conv 0.15   / 0.19     bn 0.21   / 0.27     dense 2e-3 / 4e-3
#   bf16 operands + interior  conv 0.19   / 0.22     bn 0.21   / 0.33     dense 3e-3 / 5e-3
# (WRN-16-4 at batch 32: conv 0.13, bn 0.16, dense 5e-3 in both bf16 modes.)  The bounds below leave ~1.5x head room; a wrong
# gradient gives O(1).
FIRST_FP32 = {"conv": 5e-3, "dense": 1e-3, "bn_or_bias": 1e-2}
FIRST_BF16 = {"conv": 0.30, "dense": 0.02, "bn_or_bias": 0.35}
UPD_FP32 = {"conv": 5e-2, "dense": 1e-2, "bn_or_bias": 6e-2}
UPD_BF16 = {"conv": 0.35, "dense": 0.03, "bn_or_bias": 0.50}


def test_wrn_16_4_sgd_bf16_loss_curve():
    rep = {}
    losses = _train_compare(_wrn(16, 4, 8, 16, 10), 4, db.MATH_BF16, 3e-2, None, update_tol=UPD_BF16, report=rep,
                            first_step_tol=FIRST_BF16, hyper=[H.float32((), [0.05]), H.float32((), [0.9])])
    print("update errors (bf16 operands, fp32 storage):", rep)
    assert all(np.isfinite(l[0]) for l in losses)


def test_wrn_16_4_sgd_bf16_interior_loss_curve():
    """The production configuration: bf16 tensor-core operands AND bf16 NHWC interior activations
    (DOPT_B200_PLAN_BF16_INTERIOR) against the fp32 oracle.  Stated tolerance: loss within 3e-2 relative per step,
    parameter updates within UPD_BF16."""
    rep = {}
    losses = _train_compare(_wrn(16, 4, 32, 16, 10), 4, db.MATH_BF16, 3e-2, None, flags=FUSE | GRAPH | INTERIOR,
                            update_tol=UPD_BF16, report=rep, first_step_tol=FIRST_BF16,
                            hyper=[H.float32((), [0.05]), H.float32((), [0.9])])
    print("update errors (bf16 operands, bf16 interior):", rep)
    assert all(np.isfinite(l[0]) for l in losses)


def test_bf16_interior_pass_engages_and_agrees_with_fp32_storage(monkeypatch, capfd):
    """With DOPT_B200_PLAN_BF16_INTERIOR the plan keeps the activations between tensor-core convolutions as NHWC bf16: the
    dump must show flat kernels and convolutions writing bf16, the plan must own less device memory, and two training steps
    must agree with the fp32-storage plan within bf16 rounding of the activations."""
    def run(flags):
        monkeypatch.setenv("DOPT_B200_PLAN_DUMP", "1")
        H.reset()
        H.set_math(db.MATH_BF16)
        H.set_plan_flags(flags)
        loss, extra, net, feed = _wrn(16, 4, 32, 16, 10)()   # batch 32: every residual sum is above the small-region size
        capfd.readouterr()
        upd = H.Updater(H.SGD, [loss] + extra, network=net, hyper=[H.float32((), [0.05]), H.float32((), [0.9])])
        init = [p.get().copy() for p in net.params]
        outs = [upd.step(feed(s)) for s in range(2)]
        dump = capfd.readouterr().err
        return outs, [p.get().copy() for p in net.params], upd.stats(), dump, init
    o0, p0, st0, d0, init = run(FUSE | GRAPH)
    o1, p1, st1, d1, _ = run(FUSE | GRAPH | INTERIOR)
    assert "PLAN residency" not in d0
    line = [l for l in d1.splitlines() if l.startswith("PLAN residency") and "flat kernels" in l][0]
    n_flat, n_conv, n_copies = [int(t) for t in line.replace(",", " ").split() if t.isdigit()]
    # WRN-16-4: 13 batch norms; all but the first (fp32 stem output) and the last (feeds the mean pool) run flat, forward and
    # backward, plus the residual adds
    assert n_flat >= 2 * 11 and n_conv >= 20 and n_copies <= 2, line
    assert st1["device_bytes"] < st0["device_bytes"]
    assert st1["launches"] <= st0["launches"] + 2
    for a, b in zip(o0, o1):
        assert abs(float(a[0]) - float(b[0])) <= 2e-2 * abs(float(a[0]))
        assert np.abs(a[1] - b[1]).max() <= 3e-2
    for a, b, i0 in zip(p0, p1, init):
        du = float(np.linalg.norm((a - i0).astype(np.float64)))
        if du > 1e-12:     # the update each tensor received, bf16 interior against fp32 storage
            # (two runs of the SAME bf16-interior plan differ by up to 0.11 of a tensor's update after one step -- re-quantisation
            # noise, profiles/r02_parity.md -- so 0.3 after two steps is the noise bound with head room, not a loose one)
            assert float(np.linalg.norm((a - b).astype(np.float64))) <= 0.3 * du, a.shape


def _interior_run(monkeypatch, capfd, env, steps=2, net=(16, 4, 32, 16, 10)):
    for k in ("DOPT_B200_EPI_ADD", "DOPT_B200_EPI_BNGRAD"):
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    monkeypatch.setenv("DOPT_B200_PLAN_DUMP", "1")
    H.reset()
    H.set_math(db.MATH_BF16)
    H.set_plan_flags(FUSE | GRAPH | INTERIOR)
    loss, extra, netw, feed = _wrn(*net)()
    capfd.readouterr()
    upd = H.Updater(H.SGD, [loss] + extra, network=netw, hyper=[H.float32((), [0.05]), H.float32((), [0.9])])
    init = [p.get().copy() for p in netw.params]
    outs = [upd.step(feed(s)) for s in range(steps)]
    dump = capfd.readouterr().err
    line = [l for l in dump.splitlines() if l.startswith("PLAN residency") and "epilogue" in l][0]
    n_add, n_bn = [int(t) for t in line.replace(",", " ").split() if t.isdigit()]
    return outs, [p.get().copy() for p in netw.params], init, upd.stats(), (n_add, n_bn)


def test_backward_bn_statistics_in_the_feature_gradient_epilogue(monkeypatch, capfd):
    """Plan pass H2 (opt-in, DOPT_B200_EPI_BNGRAD=1: measured slower than the separate statistics launch on the WRN-28-10
    shapes, profiles/r02_summary.md): the unit-stride convolutionFeaturesGrad that writes dy of a flat batchNormGrad accumulates sum(g) and
    sum(g * (x - mean)) in its epilogue (tc_kernel<.., EPI = 3>) from the bf16 values it stores -- the same values the
    statistics kernel would read back -- so only the fp32 summation order differs.  Stated tolerance: step-0 loss within 5e-4
    relative (the forward pass does not change; two runs of one plan differ by up to 6e-5 through the atomics of the forward
    statistics), the median parameter tensor's update within 0.12 and every tensor within 0.25 -- the measured run-to-run
    differences of the unchanged plan are 0.06 / 0.11."""
    base = {"DOPT_B200_EPI_BNGRAD": "1"}
    off = {}

    def update_differences(pa, pb, init):
        out = []
        for a, b, i0 in zip(pa, pb, init):
            du = float(np.linalg.norm((a - i0).astype(np.float64)))
            if du > 1e-12:
                out.append(float(np.linalg.norm((a - b).astype(np.float64))) / du)
        return np.array(out)
    # one step: the update is lr * gradient at identical parameters.  Two runs of the SAME bf16-interior plan already differ:
    # the forward statistics are summed with fp32 atomics, a few bf16 roundings flip, the step-0 loss lands on one of three or
    # four values 1.5e-5 apart, and the gradients of two runs that landed on different values differ by 0.06 of the update in
    # the median tensor and 0.11 in the worst (runs that landed on the same value: 1e-3; fp32 storage: 3e-6) -- measured with
    # tools/noise_probe.py, profiles/r02_parity.md: every bf16 store re-quantises, which turns a 1e-7 difference into one-ulp
    # flips, and after a few layers two runs differ by the bf16 rounding noise itself.  The bounds are those of that noise.
    o0, p0, init, st0, c0 = _interior_run(monkeypatch, capfd, off, steps=1)
    o0b, p0b, _, _, _ = _interior_run(monkeypatch, capfd, off, steps=1)
    o1, p1, _, st1, c1 = _interior_run(monkeypatch, capfd, base, steps=1)
    assert c0 == (0, 0) and c1[0] == 0 and c1[1] >= 8, (c0, c1)   # WRN-16-4: 6 blocks, 9 of 11 flat backward batch norms
    assert st1["launches"] <= st0["launches"] - c1[1]           # one statistics launch less per fused batch norm
    assert abs(float(o0[0][0]) - float(o1[0][0])) <= 5e-4 * abs(float(o0[0][0]))
    assert np.abs(o0[0][1] - o1[0][1]).max() <= 5e-3
    noise = update_differences(p0, p0b, init)
    diff = update_differences(p0, p1, init)
    print("backward statistics in the epilogue: gradient difference median %.3g worst %.3g; run-to-run noise median %.3g worst %.3g"
          % (np.median(diff), diff.max(), np.median(noise), noise.max()))
    assert np.median(diff) <= 0.12
    assert diff.max() <= 0.25


def test_residual_sum_in_the_convolution_epilogue(monkeypatch, capfd):
    """Plan pass H1 (opt-in, DOPT_B200_EPI_ADD=1: measured slower than the separate add, profiles/r02_summary.md): the second convolution of a residual block adds the block's input in its epilogue (tc_kernel<.., EPI = 2>,
    fp32 add, ONE rounding to bf16) instead of storing its bf16 result for a separate add kernel (two roundings).  Stated
    tolerance: the double rounding it removes, i.e. loss within 1e-2 relative, predictions within 3e-2, parameter updates
    within 0.3 of the update (the same bounds as bf16 interior storage against fp32 storage)."""
    o0, p0, init, st0, c0 = _interior_run(monkeypatch, capfd, {})
    o1, p1, _, st1, c1 = _interior_run(monkeypatch, capfd, {"DOPT_B200_EPI_ADD": "1"})
    assert c0 == (0, 0) and c1[1] == 0 and c1[0] >= 5, (c0, c1)   # WRN-16-4: 6 residual sums, the last one feeds a legacy kernel
    assert st1["launches"] <= st0["launches"] - c1[0]
    assert st1["device_bytes"] < st0["device_bytes"]
    for a, b in zip(o0, o1):
        assert abs(float(a[0]) - float(b[0])) <= 1e-2 * abs(float(a[0]))
        assert np.abs(a[1] - b[1]).max() <= 3e-2
    for a, b, i0 in zip(p0, p1, init):
        du = float(np.linalg.norm((a - i0).astype(np.float64)))
        if du > 1e-12:
            # (two runs of the SAME bf16-interior plan differ by up to 0.11 of a tensor's update after one step -- re-quantisation
            # noise, profiles/r02_parity.md -- so 0.3 after two steps is the noise bound with head room, not a loose one)
            assert float(np.linalg.norm((a - b).astype(np.float64))) <= 0.3 * du, a.shape


@pytest.mark.parametrize("mode", ["fp32", "bf16", "bf16-interior"])
def test_wrn_28_10_headline_graph_first_steps(mode):
    """The benchmarked graph itself (examples/cifar100.d:33-49 with wideResNet(features, 28, 10): 28 convolutions, 25 batch
    norms, 36.5 M parameters) at batch 4 for three SGD steps (lr 0.1, momentum 0.9, wd 1e-4) against the oracle: strict fp32,
    bf16 tensor-core operands, and the production configuration with bf16 interior activations."""
    def build():
        H.seed(1234)
        x, y = H.float32((4, 3, 32, 32)), H.float32((4, 100))
        preds = H.wide_resnet(x, 28, 10, weight_decay=1e-4).dense(100).softmax()
        net = H.Network([x], [preds])
        loss = H.cross_entropy(preds.train_output, y) + net.param_loss
        rng = np.random.RandomState(21)
        data = [((rng.rand(4, 3, 32, 32) * 2 - 1).astype(F), np.eye(100, dtype=F)[rng.randint(0, 100, 4)]) for _ in range(3)]
        return loss, [preds.train_output], net, lambda s: {x: data[s % 3][0], y: data[s % 3][1]}
    math = db.MATH_FP32 if mode == "fp32" else db.MATH_BF16
    flags = FUSE | GRAPH | (INTERIOR if mode == "bf16-interior" else 0)
    rep = {}
    losses = _train_compare(build, 3, math, 1e-3 if mode == "fp32" else 3e-2, None, flags=flags,
                            update_tol=UPD_FP32 if mode == "fp32" else UPD_BF16, report=rep,
                            first_step_tol=FIRST_FP32 if mode == "fp32" else FIRST_BF16,
                            hyper=[H.float32((), [0.1]), H.float32((), [0.9])])
    print("WRN-28-10 batch 4,", mode, "losses (gpu, oracle):", losses, "update errors:", rep)
    assert all(np.isfinite(l[0]) for l in losses)


def test_bn_relu_staging_absorption_is_bit_exact(monkeypatch):
    # the plan folds relu and the NHWC bf16 staging into the batch-norm apply pass (plan.cu, pass "absorb"); same arithmetic,
    # so the forward results are identical bit for bit; parameters only differ by the summation order of the fp32 atomics in
    # the filter-gradient kernel (which also varies from run to run without the fusion)
    def run(no_absorb):
        if no_absorb:
            monkeypatch.setenv("DOPT_B200_NO_ABSORB", "1")
        else:
            monkeypatch.delenv("DOPT_B200_NO_ABSORB", raising=False)
        H.reset()
        H.set_math(db.MATH_BF16)
        H.set_plan_flags(FUSE | GRAPH)
        loss, extra, net, feed = _wrn(10, 4, 8, 16, 10)()
        upd = H.Updater(H.SGD, [loss] + extra, network=net, hyper=[H.float32((), [0.05]), H.float32((), [0.9])])
        outs = [upd.step(feed(s)) for s in range(2)]
        st = upd.stats()
        return outs, [p.get().copy() for p in net.params], st["launches"]
    o0, p0, n0 = run(True)
    o1, p1, n1 = run(False)
    assert n1 < n0, (n0, n1)          # relu + staging launches are gone
    assert float(o0[0][0]) == float(o1[0][0]) and np.array_equal(o0[0][1], o1[0][1])   # step 0: loss and predictions
    assert abs(float(o0[1][0]) - float(o1[1][0])) <= 1e-5 * abs(float(o0[1][0]))
    for a, b in zip(p0, p1):
        assert float(np.abs(a - b).max()) <= 1e-5 * max(float(np.abs(a).max()), 1e-3)


@pytest.mark.parametrize("knob", ["DOPT_B200_NO_MSUM", "DOPT_B200_NO_FILTER_STAGE", "DOPT_B200_NO_BN_DIRECT", "DOPT_B200_NO_GATE",
                                  "DOPT_B200_NO_MERGE", "DOPT_B200_NO_SINGLES", "DOPT_B200_NO_BATCH", "DOPT_B200_NO_ADD_STAGE", "DOPT_B200_NO_ADD_ABSORB"])
def test_plan_passes_do_not_change_results(monkeypatch, knob):
    # every lowering pass of the plan compiler can be switched off with an environment knob (read when the plan is built):
    # batched reductions, filter staging, direct BN statistics, the relu gate in batchNormGrad, region merging, single-node
    # regions, launch batching.  Forward results must be identical bit for bit with and without each of them; the loss
    # (whose weight-decay sums change summation order under NO_MSUM) and the parameters within fp32 rounding.
    def run(off):
        if off:
            monkeypatch.setenv(knob, "1")
        else:
            monkeypatch.delenv(knob, raising=False)
        H.reset()
        H.set_math(db.MATH_BF16)
        H.set_plan_flags(FUSE | GRAPH)
        loss, extra, net, feed = _wrn(10, 4, 8, 16, 10)()
        upd = H.Updater(H.SGD, [loss] + extra, network=net, hyper=[H.float32((), [0.05]), H.float32((), [0.9])])
        outs = [upd.step(feed(s)) for s in range(2)]
        return outs, [p.get().copy() for p in net.params]
    o0, p0 = run(True)
    o1, p1 = run(False)
    assert np.array_equal(o0[0][1], o1[0][1])                                  # step-0 predictions
    assert abs(float(o0[0][0]) - float(o1[0][0])) <= 2e-6 * abs(float(o0[0][0]))
    assert abs(float(o0[1][0]) - float(o1[1][0])) <= 1e-5 * abs(float(o0[1][0]))
    for a, b in zip(p0, p1):
        assert float(np.abs(a - b).max()) <= 1e-5 * max(float(np.abs(a).max()), 1e-3)


def test_precompiled_update_program_is_bit_identical_to_the_interpreter(monkeypatch, capfd):
    """The SGD + momentum + weight-decay update of a weight tensor runs as a pre-compiled straight-line program (fused.cu,
    FzStatic<0>) instead of through the region interpreter.  Same instruction list, same dbk::apply routines: parameters
    and losses must be identical bit for bit with the table switched off.  A dense net keeps every other kernel of the
    step deterministic (the convolution filter gradients accumulate with fp32 atomics, whose order varies run to run)."""
    rng = np.random.RandomState(17)
    data = [(rng.randn(32, 64).astype(F), np.eye(10, dtype=F)[rng.randint(0, 10, 32)]) for _ in range(3)]

    def run(off):
        if off:
            monkeypatch.setenv("DOPT_B200_NO_STATIC", "1")
        else:
            monkeypatch.delenv("DOPT_B200_NO_STATIC", raising=False)
        monkeypatch.setenv("DOPT_B200_PLAN_DUMP", "1")
        H.reset()
        H.seed(18)
        H.set_math(db.MATH_FP32)
        H.set_plan_flags(FUSE | GRAPH)
        x, y = H.float32((32, 64)), H.float32((32, 10))
        l = H.data_source(x).dense(4096, weight_decay=1e-3).relu().dense(10, weight_decay=1e-3).softmax()
        net = H.Network([x], [l])
        loss = H.cross_entropy(l.train_output, y) + net.param_loss
        upd = H.Updater(H.SGD, [loss], network=net, hyper=[H.float32((), [0.05]), H.float32((), [0.9])])
        capfd.readouterr()
        outs = [float(upd.step({x: f, y: t})[0]) for f, t in data]
        dump = capfd.readouterr().err
        return outs, [p.get().copy() for p in net.params], dump
    o0, p0, d0 = run(True)
    o1, p1, d1 = run(False)
    assert "static=0" in d1 and "static=0" not in d0          # the table was really used / really off
    assert o0 == o1
    for a, b in zip(p0, p1):
        np.testing.assert_array_equal(a, b)


def test_wrn_strided_stem_sins_like_amsgrad():
    # sins10.d uses strides [2,2,2]; BASELINE configs[4] trains it with AMSGrad
    _train_compare(_wrn(10, 2, 4, 24, 10, strides=(2, 2, 2)), 2, db.MATH_FP32, 5e-4, 1e-2, kind=H.AMSGRAD,
                   hyper=[H.float32((), [1e-3]), None, None, None])


def test_vgg_with_batchnorm_sgd():
    # configs[2]: VGG-style CIFAR net with batchNorm + SGD (narrow extractor so the oracle stays fast)
    def build():
        H.seed(14)
        x, y = H.float32((4, 3, 32, 32)), H.float32((4, 10))
        l = H.data_source(x)
        for c in (8, -1, 16, -1, 16, -1, 32, -1, 32, -1):
            l = l.max_pool((2, 2)) if c == -1 else l.conv2d(c, (3, 3), padding=(1, 1)).batch_norm().relu()
        l = l.dense(32).relu().dense(10).softmax()
        net = H.Network([x], [l])
        loss = H.cross_entropy(l.train_output, y) + net.param_loss
        rng = np.random.RandomState(7)
        data = [((rng.rand(4, 3, 32, 32) * 2 - 1).astype(F), np.eye(10, dtype=F)[rng.randint(0, 10, 4)]) for _ in range(3)]
        return loss, [l.train_output], net, lambda s: {x: data[s % 3][0], y: data[s % 3][1]}
    _train_compare(build, 3, db.MATH_FP32, 5e-4, 1e-2)


def test_inference_plan_and_checkpoint_roundtrip(tmp_path):
    H.seed(15)
    x = H.float32((4, 8, 8, 8))
    l = H.data_source(x).conv2d(8, (3, 3), padding=(1, 1)).batch_norm().relu().dense(5).softmax()
    net = H.Network([x], [l])
    xs = np.random.RandomState(8).randn(4, 8, 8, 8).astype(F)
    got = H.Plan([l.output]).execute({x: xs})[0]          # test-time graph: batchNormInference
    want = G.evaluate_ops(H, [l.output], {x: xs})[0]
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5)
    path = str(tmp_path / "net.bin")
    net.save(path)
    before = [p.get() for p in net.params]
    for p in net.params:
        p.set(np.zeros(p.shape, F))
    net.load(path)
    for p, b in zip(net.params, before):
        np.testing.assert_array_equal(p.get(), b)
    import os
    assert os.path.getsize(path) == 4 * sum(p.volume for p in net.params)   # raw fp32, no header (networks.d:130-164)


@pytest.mark.parametrize("flags", [0, FUSE | GRAPH], ids=["plain", "fused+graph"])
def test_uniform_draws_fresh_numbers_every_execution(flags):
    """`uniform` (cuda/source/dopt/cuda/random.d:56-83) feeds dropout masks: every plan execution must draw new numbers, also
    when the step is replayed from a CUDA graph (the call counter lives in device memory)."""
    H.set_plan_flags(flags)
    u = H.create("uniform", [], shape=[1 << 16])
    keep = H.binary("gt", u, H.constant((1 << 16,), np.full(1 << 16, 0.25, F)))     # dropout's mask (layers/dropout.d:22-23)
    p = H.Plan([u, keep])
    draws = [p.execute() for _ in range(5)]
    for v, k in draws:
        assert v.min() > 0.0 and v.max() <= 1.0
        assert abs(float(v.mean()) - 0.5) < 1e-2
        np.testing.assert_array_equal(k, (v > 0.25).astype(F))
        assert abs(float(k.mean()) - 0.75) < 1e-2
    for i in range(len(draws)):
        for j in range(i):
            assert not np.array_equal(draws[i][0], draws[j][0])


@pytest.mark.parametrize("flags", [0, FUSE | GRAPH], ids=["plain", "fused+graph"])
def test_dropout_layer_trains_with_a_fresh_consistent_mask(flags):
    """nnet/layers/dropout.d:14-29.  Within one execution the forward mask and the mask gating the gradient are the same
    draw (one `uniform` node); across executions the mask changes; the kept fraction is 1 - dropProb."""
    H.set_plan_flags(flags)
    xs = (np.random.RandomState(9).rand(64, 512) + 1.0).astype(F)            # strictly positive: kept <=> output != 0
    x = H.float32((64, 512), xs)
    l = H.data_source(x).dropout(0.3)
    g = H.grad(H.sum_(l.train_output), [x])[0]
    p = H.Plan([l.train_output, g, l.output])
    masks = []
    for _ in range(4):
        y, gx, y_test = p.execute()
        mask = (y != 0).astype(F)
        np.testing.assert_array_equal(y, mask * xs)
        np.testing.assert_array_equal(gx, mask)
        np.testing.assert_allclose(y_test, xs * F(0.7), rtol=1e-6)
        assert abs(float(mask.mean()) - 0.7) < 0.02
        masks.append(mask)
    for i in range(len(masks)):
        for j in range(i):
            assert not np.array_equal(masks[i], masks[j])


@pytest.mark.parametrize("opts,mode", [
    (dict(dropout=True, lipschitz_norm=float("inf"), max_norm=4.0), "fp32"),
    (dict(dropout=True, maxgain_norm=2.0, max_norm=3.0, spectral_decay=1e-3), "fp32"),
    (dict(dropout=True, maxgain_norm=2.0, max_norm=3.0), "bf16-interior"),
], ids=["dropout+lipschitz-inf", "dropout+maxgain2+spectral-decay", "dropout+maxgain2/bf16-interior"])
def test_wrn_regularised_training_steps_vs_oracle(opts, mode):
    """a Wide ResNet with the regulariser options of wrn.d:11-54 -- dropout (nnet/layers/dropout.d:14-29),
    operator-norm / max-gain projections on every convolution and batch norm (nnet/lipschitz.d:28-165,
    nnet/layers/batchnorm.d:97-127), spectral decay -- trained for three SGD steps through the plan against the oracle.  The
    reference draws `uniform` unseeded, so the draws of the GPU step (dropout masks, power-iteration start vectors) are read
    back as extra plan outputs and handed to the oracle; everything else is computed independently."""
    H.set_math(db.MATH_FP32 if mode == "fp32" else db.MATH_BF16)
    H.set_plan_flags(FUSE | GRAPH | (INTERIOR if mode == "bf16-interior" else 0))
    batch = 8 if mode == "fp32" else 32
    H.seed(31)
    x, y = H.float32((batch, 3, 16, 16)), H.float32((batch, 10))
    preds = H.wide_resnet(x, 10, 2, weight_decay=1e-4, **opts).dense(10).softmax()
    net = H.Network([x], [preds])
    loss = H.cross_entropy(preds.train_output, y) + net.param_loss
    hyper = [H.float32((), [0.05]), H.float32((), [0.9])]
    probe = H.Updater(H.SGD, [loss], network=net, hyper=hyper)               # only to enumerate the graph's uniform nodes
    draws = [n["op"] for n in H.export(probe.plan_outputs()[0]) if n["type"] == "uniform"]
    assert len(draws) >= 3, "the regularised graph must contain dropout masks"
    upd = H.Updater(H.SGD, [loss, preds.train_output] + draws, network=net, hyper=hyper)
    types = [n["type"] for n in H.export(upd.plan_outputs()[0])]
    assert types.count("maxElement") >= 8 and types.count("uniform") == len(draws)
    oracle = G.UpdaterOracle(upd)
    initial = [p.get().copy() for p in net.params]
    rng = np.random.RandomState(17)
    tol = 1e-3 if mode == "fp32" else 3e-2
    losses = []
    for s in range(3):
        args = {x: (rng.rand(batch, 3, 16, 16) * 2 - 1).astype(F), y: np.eye(10, dtype=F)[rng.randint(0, 10, batch)]}
        got = upd.step(args)
        for d in got[2:]:
            assert d.min() > 0 and d.max() <= 1                              # (0, 1] like cuRAND (random.d:56-83)
        want = oracle.step(args, dict(zip(draws, got[2:])))
        losses.append((float(got[0]), float(want[0])))
        assert abs(got[0] - want[0]) <= tol * max(1.0, abs(float(want[0]))), (s, losses)
        assert np.abs(got[1] - want[1]).max() <= 10 * tol
        if s == 0:
            first = _update_errors(net.params, initial, oracle)
            for cls, (err, shape) in first.items():
                assert err <= (FIRST_FP32 if mode == "fp32" else FIRST_BF16)[cls], ("first step", cls, shape, err)
    worst = _update_errors(net.params, initial, oracle)
    print("regularised WRN-10-2,", sorted(opts), mode, "losses (gpu, oracle):", losses, "update errors:", worst)
    for cls, (err, shape) in worst.items():
        assert err <= (UPD_FP32 if mode == "fp32" else UPD_BF16)[cls], ("all steps", cls, shape, err)
    # the projections hold on the GPU's parameters: no convolution's operator-norm bound exceeds max_norm afterwards
    assert all(np.isfinite(p.get()).all() for p in net.params)


def test_lipschitz_projection_graphs_on_gpu():
    """nnet/lipschitz.d graphs (abs / axis sums / maxElement / transpose / matmul / uniform / convolutionTranspose) through
    the plan: the deterministic norms against numpy, the power-iteration norm against its bound, the projection against
    the oracle value."""
    H.set_plan_flags(FUSE | GRAPH)
    rng = np.random.RandomState(5)
    wv = rng.randn(64, 96).astype(F)
    w = H.float32((64, 96), wv)
    kv = (rng.randn(8, 4, 3, 3) * 0.2).astype(F)
    k = H.float32((8, 4, 3, 3), kv)
    n1, ninf, n2, c1, c2, proj = H.Plan([
        H.matrix_norm(w, 1), H.matrix_norm(w, float("inf")), H.matrix_norm(w, 2), H.conv_params_norm(k, (16, 16), p=1),
        H.conv_params_norm(k, (16, 16), (1, 1), (1, 1), 2),
        H.max_norm(w, H.matrix_norm(w, 1), H.float32((), [1.0]))]).execute()
    np.testing.assert_allclose(n1, np.abs(wv).sum(axis=1).max(), rtol=1e-5)
    np.testing.assert_allclose(ninf, np.abs(wv).sum(axis=0).max(), rtol=1e-5)
    true2 = np.linalg.svd(wv.astype(np.float64), compute_uv=False)[0]
    assert 0.6 * true2 <= float(n2) <= true2 * (1 + 1e-4)
    np.testing.assert_allclose(c1, np.abs(kv.reshape(8, -1)).sum(axis=1).max(), rtol=1e-5)
    bound = sum(np.linalg.svd(kv[:, :, r, q].astype(np.float64), compute_uv=False)[0] for r in range(3) for q in range(3))
    assert 0 < float(c2) <= bound * (1 + 1e-4)
    np.testing.assert_allclose(proj, wv / np.abs(wv).sum(axis=1).max(), rtol=1e-5)
