"""Host logic on CPU: the C++ mirror of dopt.core / dopt.nnet / dopt.online builds the same graphs the D code would
(checked structurally and by evaluating them with the CPU oracle against the reference's own unit-test vectors).
No GPU and no compute through the product here: the oracle is the evaluator."""
import numpy as np
import pytest

from dopt_b200 import host as H
from oracle import graph_eval as G

F = np.float32


@pytest.fixture(autouse=True)
def _fresh():
    H.init()
    H.reset()
    H.set_data_parallel_world(1)
    yield
    H.reset()


def ev(ops, args=None):
    return G.evaluate_ops(H, ops, args)


def types(ops):
    return [n["type"] for n in H.export(ops)]


# ---- reference unit tests re-run through the mirrored graph API + oracle evaluator -----------------------------------
def test_core_ops_kats():
    # core/source/dopt/core/ops/math.d:212-231
    c = H.matmul(H.float32((2, 1), [1, 2]), H.float32((1, 2), [3, 4]))
    assert ev([c])[0].ravel().tolist() == [3, 4, 6, 8]
    # math.d:285-298 (incl. the matmul-with-ones lowering of rank-2 single-axis sums)
    m = H.float32((2, 2), [0, 1, 2, 5])
    s1, s2, s3, s4 = H.sum_(H.float32((2,), [0.5, 1.5])), H.sum_(m), H.sum_(m, [0]), H.sum_(m, [1])
    r = ev([s1, s2, s3, s4])
    assert [x.ravel().tolist() for x in r] == [[2.0], [8.0], [2.0, 6.0], [1.0, 7.0]]
    assert "matmul" in types([s3]) and "sum" not in types([s3])
    assert "sum" in types([s2])
    # math.d:318-331, 361-373
    assert ev([H.argmin(H.float32((5,), [4, 2, 6, 1, 2]), 0)])[0].ravel().tolist() == [3]
    a = H.float32((2, 2), [1, 4, 3, 6])
    assert [x.ravel().tolist() for x in ev([H.max_element(a), H.max_element(a, [0]), H.max_element(a, [1])])] == \
        [[6.0], [3.0, 6.0], [4.0, 6.0]]


def test_basic_ops_kats():
    # core/source/dopt/core/ops/basic.d:212-395
    s = H.slice_(H.int32((3, 3), list(range(1, 10))), [1, 1], [3, 3])
    assert ev([s])[0].ravel().tolist() == [5, 6, 8, 9]
    p = H.pad(H.int32((1, 1), [3]), [2, 1], [3, 3])
    got = ev([p])[0]
    assert got.shape == (6, 5) and got[2, 1] == 3 and got.sum() == 3
    r = H.reshape(H.float32((2, 2), [1, 2, 3, 4]), [1, 4])
    assert r.shape == (1, 4) and ev([r])[0].ravel().tolist() == [1, 2, 3, 4]
    t = H.transpose(H.float32((2, 2), [1, 2, 3, 4]), [1, 0])
    assert ev([t])[0].ravel().tolist() == [1, 3, 2, 4]
    r2 = H.repeat(H.float32((2, 2), [1, 2, 3, 4]), [3, 2])
    assert ev([r2])[0].ravel().tolist() == [1, 2, 1, 2, 3, 4, 3, 4] * 3
    r3 = H.repeat(H.float32((2,), [1, 2]), 3)
    assert r3.shape == (3, 2) and ev([r3])[0].ravel().tolist() == [1, 2, 1, 2, 1, 2]
    assert types([r3]).count("matmul") == 1   # repeat(n) is a matmul with a ones column (basic.d:370-381)


def test_verifiers_reject_like_the_reference():
    with pytest.raises(H.HostError):
        H.matmul(H.float32((2, 3)), H.float32((2, 3)))            # inner dimensions differ
    with pytest.raises(H.HostError):
        H.binary("add", H.float32((2,)), H.float32((3,)))         # pointwise operands must have identical types
    with pytest.raises(H.HostError):
        H.slice_(H.float32((3,)), [2], [2])                       # start < stop
    with pytest.raises(H.HostError):
        H.reshape(H.float32((3,)), [2, 2])
    with pytest.raises(H.HostError):
        H.create("noSuchOp", [])
    with pytest.raises(H.HostError):
        H.grad(H.float32((2,)), [H.float32((2,))])                # objective must have volume one


def test_nnet_ops_kats():
    # core/source/dopt/core/ops/nnet.d:270-293, 333-352, 424-435
    y = H.convolution(H.float32((1, 1, 3, 5), [1, 1, 1, 0, 0] * 3), H.float32((1, 1, 1, 2), [-1, 1]))
    assert y.shape == (1, 1, 3, 4) and ev([y])[0].ravel().tolist() == [0, 0, 1, 0] * 3
    mp = H.maxpool(H.float32((1, 1, 4, 4), [1, 2, 4, 3, 5, 3, 2, 2, 0.1, -4, 3, 2, 0, 0, 2, 2]), [2, 2])
    assert ev([mp])[0].ravel().tolist() == [F(5), F(4), F(0.1), F(3)]
    sm = H.softmax(H.float32((1, 5), [1, 2, 3, 1, 2]))
    np.testing.assert_allclose(ev([sm])[0].ravel(), [0.0674508, 0.18335, 0.498398, 0.0674508, 0.18335], atol=1e-6)
    # convolutionTranspose's shape rule (nnet.d:305-315)
    ct = H.convolution_transpose(H.float32((2, 8, 5, 5)), H.float32((8, 3, 3, 3)), (1, 1), (2, 2))
    assert ct.shape == (2, 3, 9, 9)


def test_scalar_broadcast_lowering():
    # Operation.opBinary (ops/package.d:96-103): rank-0 operand -> repeat(volume).reshape(shape) -> matmul with ones
    x = H.float32((2, 3), np.arange(6))
    y = x * 2.0
    assert ev([y])[0].ravel().tolist() == [0, 2, 4, 6, 8, 10]
    ts = types([y])
    assert ts.count("matmul") == 1 and ts.count("constant") == 2 and ts[-1] == "mul"
    z = 1.0 - x
    assert ev([z])[0].ravel().tolist() == [1, 0, -1, -2, -3, -4]


def test_autodiff_kats():
    # core/source/dopt/core/grads/package.d:113-129: d(sum(x*x))/dx == 2x
    rng = np.random.RandomState(0)
    xv = rng.randn(3, 4).astype(F)
    x = H.float32((3, 4), xv)
    g = H.grad(H.sum_(x * x), [x])[0]
    np.testing.assert_allclose(ev([g])[0], 2 * xv, rtol=1e-6)
    # core/source/dopt/core/grads/basic.d:49-67: sliceGrad -> pad
    a = H.float32((4, 4), [1, 2, 3, 4] * 4)
    b = H.float32((4, 4), [5, 6, 7, 8] * 4)
    c = H.slice_(a * b, [1, 1], [2, 2])
    ga = H.grad(c, [a])[0]
    assert ev([ga])[0].ravel().tolist() == [0, 0, 0, 0, 0, 6, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]
    assert "pad" in types([ga])


def test_gradient_graph_matches_finite_differences_for_a_small_net():
    rng = np.random.RandomState(1)
    H.seed(3)
    x = H.float32((4, 3, 6, 6))
    labels = H.float32((4, 5))
    out = H.data_source(x).conv2d(4, (3, 3), padding=(1, 1), weight_decay=1e-2).batch_norm().relu().max_pool((2, 2)) \
        .dense(5).softmax()
    net = H.Network([x], [out])
    loss = H.cross_entropy(out.train_output, labels) + net.param_loss
    params = net.params
    grads = H.grad(loss, params[:2])   # conv filters and conv bias
    xs = rng.randn(4, 3, 6, 6).astype(F)
    ls = np.eye(5, dtype=F)[rng.randint(0, 5, 4)]
    base = ev([loss] + grads, {x: xs, labels: ls})
    w0 = params[0].get()
    eps = 1e-2
    for idx in [(0, 0, 0, 0), (3, 2, 1, 2), (1, 1, 2, 0)]:
        wp, wm = w0.copy(), w0.copy()
        wp[idx] += eps
        wm[idx] -= eps
        params[0].set(wp)
        lp = ev([loss], {x: xs, labels: ls})[0]
        params[0].set(wm)
        lm = ev([loss], {x: xs, labels: ls})[0]
        params[0].set(w0)
        fd = (float(lp) - float(lm)) / (2 * eps)
        assert abs(fd - base[1][idx]) < 5e-3 + 5e-2 * abs(fd), (idx, fd, base[1][idx])


def test_batchnorm_layer_structure_and_running_mean_kat():
    # nnet/source/dopt/nnet/layers/batchnorm.d:158-177 with dopt.online.adam and the mean/var "projections"
    x = H.float32((3, 2), [1, 2, 3, 4, 5, 6])
    layer = H.data_source(x).batch_norm()
    net = H.Network([x], [layer])
    params = net.params
    assert [p.shape for p in params] == [(1, 2, 1, 1), (2,), (2,), (2,)]
    trloss = H.sum_(layer.train_output)
    upd = H.Updater(H.ADAM, [trloss], network=net)
    plan_ops, dests = upd.plan_outputs()
    # outputs ~ newvals(4) ~ means(4) ~ vars(4) ~ [nb1, nb2]
    assert len(plan_ops) == 1 + 4 + 4 + 4 + 2
    ts = types(plan_ops)
    assert ts.count("batchNormTrain") == 1 and ts.count("batchNormGrad") == 1
    oracle = G.UpdaterOracle(upd)
    for _ in range(1000):
        oracle.step({})
    np.testing.assert_allclose(oracle.value_of(params[2]), [3.0, 4.0], rtol=1e-2, atol=1e-5)


@pytest.mark.parametrize("kind", [H.SGD, H.ADAM, H.AMSGRAD])
def test_optimisers_fit_a_line(kind):
    # online/source/dopt/online/{sgd.d:97-141, adam.d:96-140, amsgrad.d:102-146}: fit y = 3x + 2 (the reference only prints)
    rng = np.random.RandomState(2)
    xdata = rng.rand(100).astype(F)
    ydata = 3.0 * xdata + 2.0
    m, c = H.float32((), [0.0]), H.float32((), [0.0])
    x, y = H.float32((100,)), H.float32((100,))
    yhat = m * x + c
    diff = yhat - y
    loss = H.sum_(diff * diff) * (1.0 / 100)
    if kind == H.SGD:
        upd = H.Updater(kind, [loss], wrt=[m, c], hyper=[H.float32((), [0.1]), H.float32((), [0.5])])
        steps = 600
    else:
        upd = H.Updater(kind, [loss], wrt=[m, c], hyper=[H.float32((), [0.1]), None, None, None])
        steps = 600
    oracle = G.UpdaterOracle(upd)
    first = last = None
    for i in range(steps):
        out = oracle.step({x: xdata, y: ydata})[0]
        first = out if first is None else first
        last = out
    assert float(last) < 1e-3 < float(first)
    assert abs(float(oracle.value_of(m)) - 3.0) < 0.1 and abs(float(oracle.value_of(c)) - 2.0) < 0.1


def test_sgd_graph_shape_and_amsgrad_quirk():
    w = H.float32((5,), np.arange(5))
    loss = H.sum_(w * w)
    upd = H.Updater(H.SGD, [loss], wrt=[w])
    plan_ops, dests = upd.plan_outputs()
    assert len(plan_ops) == 3 and dests[0] is None and dests[1].serial == w.serial
    # m' = m*mu + lr*g ; w' = w - m' : mul, mul, add, sub after the gradient (sgd.d:57-64)
    tail = [t for t in types(plan_ops) if t in ("add", "sub", "mul")]
    assert tail.count("sub") == 1
    upd = H.Updater(H.AMSGRAD, [loss], wrt=[w])
    plan_ops, dests = upd.plan_outputs()
    assert len(plan_ops) == 1 + 4 + 2
    # varhat' = max(varhat, OLD var) and is never read by the update (survey F11): after one step it is still zero
    oracle = G.UpdaterOracle(upd)
    oracle.step({})
    assert not oracle.value_of(dests[4]).any()


def test_wrn_graph_inventory():
    """Wide ResNet 16-4 (examples/cifar100.d:40-48): the op inventory the kernels see."""
    H.seed(1)
    x = H.float32((2, 3, 32, 32))
    labels = H.float32((2, 100))
    preds = H.wide_resnet(x, 16, 4).dense(100).softmax()
    net = H.Network([x], [preds])
    # 13 convs (1 stem + 3 blocks x (2x2 + shortcut)) no bias, 13 BNs (4 params each), dense weight + bias
    convs = [p for p in net.params if len(p.shape) == 4 and p.shape[0] != 1]
    assert len(convs) == 1 + 3 * (2 * 2 + 1)
    assert len(net.params) == len(convs) + 13 * 4 + 2
    loss = H.cross_entropy(preds.train_output, labels) + net.param_loss
    upd = H.Updater(H.SGD, [loss, preds.train_output], network=net,
                    hyper=[H.float32((), [0.1]), H.float32((), [0.9])])
    plan_ops, dests = upd.plan_outputs()
    assert len(plan_ops) == 2 + 2 * len(net.params)
    ts = types(plan_ops)
    assert ts.count("convolution") == len(convs)
    assert ts.count("convolutionFiltersGrad") == len(convs)
    assert ts.count("convolutionFeaturesGrad") == len(convs) - 1     # the stem's features are the network input
    assert ts.count("batchNormTrain") == 13 and ts.count("batchNormGrad") == 13
    assert ts.count("relu") == 13 and ts.count("reluGrad") == 13
    assert ts.count("sum") == len(convs) + 1                          # weight decay per conv + the cross-entropy sum


def test_wrn_28_10_headline_graph_inventory():
    """The benchmarked graph (BASELINE configs[3]: wideResNet(features, 28, 10).dense(100).softmax(), SGD) pinned against
    nnet/source/dopt/nnet/models/wrn.d:56-199 independently of the mirror: the expected counts and sizes below are derived
    from the D source's loops (n = (depth - 4) / 6 blocks per group, a 1x1 shortcut only where the channel count changes,
    batchNorm before both 3x3 convolutions of a block and once after the last group, every batchNorm holding scale, bias,
    mean, var)."""
    depth, width, classes = 28, 10, 100
    n = (depth - 4) // 6
    groups = [(16, 16 * width), (16 * width, 32 * width), (32 * width, 64 * width)]
    want_convs = [(16, 3, 3, 3)]
    bn_channels = []
    for cin, u in groups:
        c = cin
        for _ in range(n):
            bn_channels += [c, u]
            want_convs += [(u, c, 3, 3), (u, u, 3, 3)]
            if c != u:
                want_convs.append((u, c, 1, 1))
            c = u
    bn_channels.append(64 * width)
    assert len(want_convs) == 28 and len(bn_channels) == 25
    want_elems = sum(int(np.prod(s)) for s in want_convs) + 4 * sum(bn_channels) + 64 * width * classes + classes
    assert want_elems == 36554836

    H.seed(1)
    x = H.float32((2, 3, 32, 32))
    labels = H.float32((2, classes))
    preds = H.wide_resnet(x, depth, width, weight_decay=1e-4).dense(classes).softmax()
    net = H.Network([x], [preds])
    assert len(net.params) == 130
    assert sum(p.volume for p in net.params) == want_elems
    got_convs = sorted(tuple(p.shape) for p in net.params if len(p.shape) == 4 and p.shape[0] != 1)
    assert got_convs == sorted(want_convs)
    assert sorted(p.shape[1] for p in net.params if len(p.shape) == 4 and p.shape[0] == 1) == sorted(bn_channels)   # scale [1,C,1,1]
    loss = H.cross_entropy(preds.train_output, labels) + net.param_loss
    upd = H.Updater(H.SGD, [loss, preds.train_output], network=net, hyper=[H.float32((), [0.1]), H.float32((), [0.9])])
    plan_ops, _ = upd.plan_outputs()
    nodes = H.export(plan_ops)
    ts = [nd["type"] for nd in nodes]
    assert ts.count("convolution") == 28 and ts.count("convolutionFiltersGrad") == 28
    assert ts.count("convolutionFeaturesGrad") == 27          # the stem's features are the network input
    assert ts.count("batchNormTrain") == 25 and ts.count("batchNormGrad") == 25
    assert ts.count("relu") == 25 and ts.count("reluGrad") == 25
    # forward residual sums: rank-4 adds that the loss depends on
    by_id = dict((nd["id"], nd) for nd in nodes)
    loss_id = [nd["id"] for nd in nodes if nd["op"].h == plan_ops[0].h][0]
    fwd, stack = set(), [loss_id]
    while stack:
        i = stack.pop()
        if i not in fwd:
            fwd.add(i)
            stack.extend(by_id[i]["deps"])
    assert sum(1 for nd in nodes if nd["type"] == "add" and len(nd["shape"]) == 4 and nd["id"] in fwd) == 12
    # convolution attributes of the three strided 3x3 / 1x1 pairs (wrn.d:143-151,166-178 with stride [1,2,2])
    strides = sorted(tuple(nd["attrs"]["stride"]) for nd in nodes if nd["type"] == "convolution")
    assert strides.count((2, 2)) == 4 and strides.count((1, 1)) == 24


def test_vgg19_batchnorm_cifar10_graph_inventory():
    """BASELINE configs[2] -- vgg19(features, [512, 512]).dense(10).softmax() with batchnorm on 100 x 3 x 32 x 32
    (examples/cifar10.d:54-66) -- pinned against nnet/source/dopt/nnet/models/vgg.d:67-75,104-134 independently of the mirror:
    the expected inventory is derived from the D source's size list (conv3x3 pad 1 with bias -> batchNorm -> relu per entry,
    maxPool [2,2] per -1, then dense + relu per top size)."""
    sizes = [64, 64, -1, 128, 128, -1, 256, 256, 256, 256, -1, 512, 512, 512, 512, -1, 512, 512, 512, 512, -1]
    want_convs, cin, elems = [], 3, 0
    for sz in sizes:
        if sz == -1:
            continue
        want_convs.append((sz, cin, 3, 3))
        elems += sz * cin * 9 + sz + 4 * sz          # filter, bias, batchNorm scale / bias / mean / var
        cin = sz
    elems += (512 * 512 + 512) * 2 + 10 * 512 + 10
    assert len(want_convs) == 16 and elems == 20576842
    H.seed(1)
    x, labels = H.float32((100, 3, 32, 32)), H.float32((100, 10))
    preds = H.vgg19(x, dense_sizes=(512, 512), batchnorm=True).dense(10).softmax()
    net = H.Network([x], [preds])
    assert sum(p.volume for p in net.params) == elems
    assert sorted(tuple(p.shape) for p in net.params if len(p.shape) == 4 and p.shape[0] != 1) == sorted(want_convs)
    assert sorted(tuple(p.shape) for p in net.params if len(p.shape) == 2) == [(10, 512), (512, 512), (512, 512)]
    loss = H.cross_entropy(preds.train_output, labels) + net.param_loss
    upd = H.Updater(H.SGD, [loss, preds.train_output], network=net, hyper=[H.float32((), [0.01]), H.float32((), [0.9])])
    nodes = H.export(upd.plan_outputs()[0])
    ts = [nd["type"] for nd in nodes]
    assert ts.count("convolution") == 16 and ts.count("convolutionFiltersGrad") == 16 and ts.count("convolutionFeaturesGrad") == 15
    assert ts.count("batchNormTrain") == 16 and ts.count("batchNormGrad") == 16
    assert ts.count("maxpool") == 5 and ts.count("maxpoolGrad") == 5
    # (a dense layer adds its bias as `y + bias.repeat(N)`, nnet/layers/dense.d:143; only the convolutions use addBias)
    assert ts.count("addBias") == 16 and ts.count("addBiasGrad") == 16 and ts.count("relu") == 16 + 2
    pools = sorted(tuple(nd["shape"][2:]) for nd in nodes if nd["type"] == "maxpool")
    assert pools == [(1, 1), (2, 2), (4, 4), (8, 8), (16, 16)]        # 32 x 32 halves five times: the last maps are 2 x 2


def test_sins_wrn_16_8_graph_inventory():
    """BASELINE configs[4] -- wideResNet(features, 16, 8, stride [2,2,2]).dense(10).softmax() on 50 x 3 x 96 x 96
    (examples/sins10.d:42-52) -- against nnet/source/dopt/nnet/models/wrn.d:56-199: two blocks per group, a 1x1 shortcut where the
    channel count changes, every group's first block strided, so the maps are 48 x 48, 24 x 24 and 12 x 12."""
    depth, width, classes = 16, 8, 10
    n = (depth - 4) // 6
    want_convs, bn_channels = [(16, 3, 3, 3)], []
    for cin, u in [(16, 16 * width), (16 * width, 32 * width), (32 * width, 64 * width)]:
        c = cin
        for _ in range(n):
            bn_channels += [c, u]
            want_convs += [(u, c, 3, 3), (u, u, 3, 3)]
            if c != u:
                want_convs.append((u, c, 1, 1))
            c = u
    bn_channels.append(64 * width)
    elems = sum(int(np.prod(sh)) for sh in want_convs) + 4 * sum(bn_channels) + 64 * width * classes + classes
    assert len(want_convs) == 16 and len(bn_channels) == 13 and elems == 10968570
    H.seed(1)
    x = H.float32((50, 3, 96, 96))
    preds = H.wide_resnet(x, depth, width, stride=(2, 2, 2)).dense(classes).softmax()
    net = H.Network([x], [preds])
    assert sum(p.volume for p in net.params) == elems
    assert sorted(tuple(p.shape) for p in net.params if len(p.shape) == 4 and p.shape[0] != 1) == sorted(want_convs)
    nodes = H.export([preds.train_output])
    convs = [nd for nd in nodes if nd["type"] == "convolution"]
    assert sorted(tuple(nd["attrs"]["stride"]) for nd in convs).count((2, 2)) == 6        # three 3x3 + three 1x1 shortcuts
    assert sorted(set(tuple(nd["shape"][2:]) for nd in convs)) == [(12, 12), (24, 24), (48, 48), (96, 96)]


def test_dropout_layer_graph():
    """nnet/layers/dropout.d:14-29: train output = (uniform > p) * x, test output = x * (1 - p); the mask is not
    differentiable, so the gradient wrt x is parentGrad * mask."""
    x = H.float32((4, 6), np.arange(24))
    l = H.data_source(x).dropout(0.25)
    tr = H.export([l.train_output])
    assert [n["type"] for n in tr] == ["uniform", "variable", "gt", "variable", "mul"]
    assert tr[0]["attrs"]["shape"] == [4, 6]
    np.testing.assert_array_equal(tr[1]["op"].get(), np.full((4, 6), 0.25, F))
    np.testing.assert_allclose(ev([l.output])[0], np.arange(24, dtype=F).reshape(4, 6) * F(0.75))
    g = H.grad(H.sum_(l.train_output), [x])[0]
    ts = types([g])
    assert "gt" in ts and "uniform" in ts and ts.count("uniform") == 1       # the SAME mask node gates the gradient


@pytest.mark.parametrize("kind", [H.SGD, H.ADAM, H.AMSGRAD])
def test_training_checkpoint_roundtrip(tmp_path, kind):
    """Parameters + optimiser state (+ Adam's running b1 / b2) survive save_state / load_state; the body of the file starts
    with exactly the bytes DAGNetwork.save writes (nnet/networks.d:130-164); mismatching or truncated files are refused
    without touching the model."""
    import os
    import struct
    H.seed(3)
    x, y = H.float32((4, 3, 8, 8)), H.float32((4, 5))
    l = H.data_source(x).conv2d(4, (3, 3), padding=(1, 1), weight_decay=1e-3).batch_norm().relu().dense(5).softmax()
    net = H.Network([x], [l])
    loss = H.cross_entropy(l.train_output, y) + net.param_loss
    upd = H.Updater(kind, [loss], network=net)
    _, dests = upd.plan_outputs()
    state = [d for d in dests if d is not None]
    n_slots = {H.SGD: 2, H.ADAM: 3, H.AMSGRAD: 4}[kind]
    assert len(state) == n_slots * len(net.params) + (0 if kind == H.SGD else 2)
    rng = np.random.RandomState(4)
    want = []
    for v in state:                                              # pretend some training happened
        a = rng.randn(*v.shape).astype(F) if v.shape else np.array(rng.rand(), F)
        v.set(a)
        want.append(a)
    path, ppath = str(tmp_path / "train.ckpt"), str(tmp_path / "params.bin")
    upd.save_state(path)
    net.save(ppath)
    raw = open(path, "rb").read()
    hdr = upd.state_header_bytes()
    assert raw[:8] == b"DOPTCKPT" and struct.unpack("<II", raw[8:16]) == (1, len(state))
    assert list(struct.unpack("<%dQ" % len(state), raw[16:hdr])) == [v.volume for v in state]
    assert len(raw) == hdr + 4 * sum(v.volume for v in state)
    params = open(ppath, "rb").read()
    assert raw[hdr:hdr + len(params)] == params                  # reference-format parameter file is a prefix of the body
    for v in state:
        v.set(np.zeros(v.shape, F))
    upd.load_state(path)
    for v, a in zip(state, want):
        np.testing.assert_array_equal(v.get(), a)
    # refused: truncated file, wrong magic, an updater with a different tensor list
    open(path, "wb").write(raw[:-8])
    for v in state:
        v.set(np.ones(v.shape, F))
    with pytest.raises(H.HostError):
        upd.load_state(path)
    for v in state:
        np.testing.assert_array_equal(v.get(), np.ones(v.shape, F))   # nothing was half-restored
    open(path, "wb").write(b"NOTACKPT" + raw[8:])
    with pytest.raises(H.HostError):
        upd.load_state(path)
    open(path, "wb").write(raw)
    other = H.Updater(H.SGD if kind != H.SGD else H.ADAM, [loss], network=net)
    with pytest.raises(H.HostError):
        other.load_state(path)
    assert os.path.getsize(ppath) == 4 * sum(p.volume for p in net.params)


def test_lipschitz_norms_and_projection():
    """nnet/lipschitz.d: operator norms of a weight matrix / a convolution and the max-norm projection, evaluated with the
    oracle.  p = 1 / infinity are exact row / column sums; p = 2 is a two-step power iteration from a random start, so it
    is a lower bound that gets close to the true spectral norm."""
    rng = np.random.RandomState(5)
    wv = rng.randn(6, 9).astype(F)
    w = H.float32((6, 9), wv)
    n1, ninf, n2 = ev([H.matrix_norm(w, 1), H.matrix_norm(w, float("inf")), H.matrix_norm(w, 2)])
    np.testing.assert_allclose(n1, np.abs(wv).sum(axis=1).max(), rtol=1e-6)
    np.testing.assert_allclose(ninf, np.abs(wv).sum(axis=0).max(), rtol=1e-6)
    true2 = np.linalg.svd(wv.astype(np.float64), compute_uv=False)[0]
    assert 0.7 * true2 <= float(n2) <= true2 * (1 + 1e-5)
    assert types([H.matrix_norm(w, 2)]).count("uniform") == 1
    # convolution: p = 1 / infinity reduce to the reshaped matrix (lipschitz.d:130-138); p = 2 power-iterates conv^T conv
    kv = rng.randn(4, 3, 3, 3).astype(F)
    k = H.float32((4, 3, 3, 3), kv)
    c1 = ev([H.conv_params_norm(k, (8, 8), p=1)])[0]
    np.testing.assert_allclose(c1, np.abs(kv.reshape(4, -1)).sum(axis=1).max(), rtol=1e-6)
    ts = types([H.conv_params_norm(k, (8, 8), (1, 1), (1, 1), 2)])
    assert ts.count("convolution") == 3 and ts.count("convolutionFeaturesGrad") == 2      # n = 2 round trips + the final conv
    c2 = float(ev([H.conv_params_norm(k, (8, 8), (1, 1), (1, 1), 2)])[0])
    # the operator norm of a padded convolution is bounded by the sum over taps of the tap matrices' spectral norms
    bound = sum(np.linalg.svd(kv[:, :, r, q].astype(np.float64), compute_uv=False)[0] for r in range(3) for q in range(3))
    assert 0 < c2 <= bound * (1 + 1e-5)
    # maxNorm: scale down only when the norm exceeds the bound (lipschitz.d:162-165)
    for bound_v, factor in ((1.0, 1.0 / float(n1)), (1e6, 1.0)):
        proj = H.max_norm(w, H.matrix_norm(w, 1), H.float32((), [bound_v]))
        np.testing.assert_allclose(ev([proj])[0], wv * F(factor), rtol=2e-6)
    with pytest.raises(H.HostError):
        H.matrix_norm(k, 1)                                           # "This function only operates on matrices"


def test_maxgain_projection_bounds_the_layer_gain():
    """conv.d:125-150 / dense.d:105-131: with `maxgain` set, the projected weights are the updated weights scaled by
    1 / max(1, g / maxgain), g = max_n ||y_n|| / ||x_n|| over the train batch (evaluated with the weights BEFORE the step)."""
    rng = np.random.RandomState(6)
    xs = rng.randn(5, 3, 6, 6).astype(F)
    x = H.float32((5, 3, 6, 6), xs)
    H.seed(7)
    l = H.data_source(x).conv2d(4, (3, 3), padding=(1, 1), use_bias=False, maxgain=0.5)
    net = H.Network([x], [l])
    w = net.params[0]
    upd = H.Updater(H.SGD, [H.sum_(l.train_output * l.train_output)], network=net,
                    hyper=[H.float32((), [0.0]), H.float32((), [0.0])])     # lr 0: the step is the projection alone
    o = G.UpdaterOracle(upd)
    w0 = w.get().copy()
    o.step({})
    y = ev([l.train_output])[0]
    gain = np.sqrt(((y.reshape(5, -1) ** 2).sum(1) + 1e-8) / ((xs.reshape(5, -1) ** 2).sum(1) + 1e-8)).max()
    assert gain > 0.5                                             # the constraint is active for He-initialised filters
    np.testing.assert_allclose(o.value_of(w), w0 / (gain / 0.5), rtol=1e-5)
    # dense: same rule
    H.reset()
    xd = H.float32((6, 10), rng.randn(6, 10).astype(F))
    H.seed(8)
    ld = H.data_source(xd).dense(7, use_bias=False, maxgain=1e6)
    netd = H.Network([xd], [ld])
    updd = H.Updater(H.SGD, [H.sum_(ld.train_output)], network=netd, hyper=[H.float32((), [0.0]), H.float32((), [0.0])])
    od = G.UpdaterOracle(updd)
    wd0 = netd.params[0].get().copy()
    od.step({})
    np.testing.assert_array_equal(od.value_of(netd.params[0]), wd0)   # gain far below the bound: weights untouched


def test_batchnorm_lipschitz_projection_and_spectral_decay_graphs():
    rng = np.random.RandomState(9)
    x = H.float32((8, 3, 4, 4), (rng.randn(8, 3, 4, 4) * 2).astype(F))
    l = H.data_source(x).batch_norm(lipschitz=0.25)
    net = H.Network([x], [l])
    gamma = net.params[0]
    upd = H.Updater(H.SGD, [H.sum_(l.train_output)], network=net, hyper=[H.float32((), [0.0]), H.float32((), [0.0])])
    o = G.UpdaterOracle(upd)
    o.step({})
    # batchnorm.d:115-129: norm = max_c |gamma_c| / sqrt(newRunningVar_c + 1e-6); gamma scaled by 1 / max(1, norm / bound)
    new_var = o.value_of(net.params[3])
    norm = np.abs(np.ones(3, F) / np.sqrt(new_var + 1e-6)).max()
    np.testing.assert_allclose(o.value_of(gamma).ravel(), np.ones(3, F) / max(1.0, norm / 0.25), rtol=1e-5)
    # spectral decay adds spectralDecay * ||W v||^2 (one power iteration from a uniform start) to the parameter loss
    H.reset()
    xc = H.float32((2, 3, 5, 5))
    H.seed(10)
    lc = H.data_source(xc).conv2d(4, (3, 3), padding=(1, 1), use_bias=False, spectral_decay=0.1).dense(3, spectral_decay=0.2)
    netc = H.Network([xc], [lc])
    ts = types([netc.param_loss])
    assert ts.count("uniform") == 2
    wv = netc.params[0].get().reshape(4, -1).astype(np.float64)
    dv = netc.params[1 if len(netc.params[1].shape) == 2 else 2].get().astype(np.float64)
    loss = float(ev([netc.param_loss])[0])
    top = 0.1 * np.linalg.svd(wv, compute_uv=False)[0] ** 2 + 0.2 * np.linalg.svd(dv, compute_uv=False)[0] ** 2
    assert 0 < loss <= top * (1 + 1e-5)                            # a power-iteration estimate never exceeds the true norm


def test_wrn_regulariser_options():
    """wrn.d:11-54,104-201: dropout after the first conv-bn-relu of each block; a Lipschitz constraint puts an operator-norm
    projection on every convolution and bounds gamma / sqrt(var) in every batch norm; only one of maxgain / lipschitz."""
    x = H.float32((2, 3, 8, 8))
    y = H.float32((2, 10))

    def plan_types(**kw):
        H.seed(11)
        l = H.wide_resnet(x, 10, 1, **kw).dense(10).softmax()
        net = H.Network([x], [l])
        upd = H.Updater(H.SGD, [H.cross_entropy(l.train_output, y) + net.param_loss], network=net)
        return types(upd.plan_outputs()[0])
    base = plan_types()
    assert base.count("uniform") == 0 and base.count("maxElement") == 0
    drop = plan_types(dropout=True)
    assert drop.count("uniform") == 3                              # one per block (n = 1, three groups)
    lip = plan_types(lipschitz_norm=float("inf"), max_norm=4.0)
    # stem + 2 convs per block + a shortcut conv where the width changes (not in the first group at width 1: 16 -> 16);
    # 2 batch norms per block + the final one
    n_conv, n_bn = 1 + 3 * 2 + 2, 3 * 2 + 1
    assert lip.count("maxElement") == n_conv + n_bn
    mg = plan_types(maxgain_norm=2.0, max_norm=3.0)
    assert mg.count("maxElement") == n_conv + n_bn
    with pytest.raises(H.HostError):
        plan_types(maxgain_norm=1.0, max_norm=3.0)                 # "Only a maxgainNorm of 2 is currently supported."
    with pytest.raises(H.HostError):
        plan_types(maxgain_norm=2.0, lipschitz_norm=2.0, max_norm=3.0)


def test_vgg_regulariser_options():
    """vgg.d:12-174: dropout 0.2 before every convolution after the first pool and 0.5 before every dense layer; the
    Lipschitz option projects every convolution (projConvParams) and dense layer (projMatrix) and bounds every batch norm."""
    x = H.float32((2, 3, 32, 32))
    y = H.float32((2, 10))

    def plan_types(layers, **kw):
        H.seed(12)
        l = H.vgg(x, layers, dense_sizes=(16, 16), **kw).dense(10).softmax()
        net = H.Network([x], [l])
        upd = H.Updater(H.SGD, [H.cross_entropy(l.train_output, y)], network=net)
        return types(upd.plan_outputs()[0])
    base16, base19 = plan_types(16), plan_types(19)
    assert base16.count("convolution") == 13 and base19.count("convolution") == 16
    assert base19.count("maxpool") == 5 and base19.count("uniform") == 0
    drop = plan_types(19, dropout=True)
    assert drop.count("uniform") == (16 - 2) + 2                   # all convs but the two before the first pool + two dense layers
    lip = plan_types(19, batchnorm=True, lipschitz_norm=1.0, max_norm=5.0)
    assert lip.count("maxElement") == 16 + 16 + 2                   # 16 convs + 16 batch norms + the two regularised dense layers
    with pytest.raises(H.HostError):
        plan_types(19, maxgain_norm=2.0, lipschitz_norm=2.0)


def test_data_parallel_wraps_gradients_in_allreduce():
    w = H.float32((4,), np.arange(4))
    x = H.float32((4,), np.arange(4))          # fed by the caller: the rank's share of the data
    loss = H.sum_(w * x)
    H.set_data_parallel_world(8)
    upd = H.Updater(H.SGD, [loss], wrt=[w])
    assert types(upd.plan_outputs()[0]).count("allreduce") == 1
    # a loss that depends on the parameters alone (a pure regulariser) has the same gradient on every rank: nothing to exchange
    upd = H.Updater(H.SGD, [H.sum_(w * w)], wrt=[w])
    assert types(upd.plan_outputs()[0]).count("allreduce") == 0
    # data term + regulariser: only the data term's gradient goes through the collective
    upd = H.Updater(H.SGD, [H.sum_(w * x) + H.sum_(w * w)], wrt=[w])
    assert types(upd.plan_outputs()[0]).count("allreduce") == 1
    H.set_data_parallel_world(1)
    upd = H.Updater(H.SGD, [loss], wrt=[w])
    assert types(upd.plan_outputs()[0]).count("allreduce") == 0


def test_data_parallel_batchnorm_statistics_are_averaged_and_zero_gradients_are_not_exchanged():
    """A batch-norm layer has two trainable tensors (scale, bias) and two state tensors (running mean / var) whose "gradient"
    is a zero variable (core/source/dopt/core/grads/nnet.d:80-81) and whose new value is a projection computed from the rank's
    own batch (nnet/layers/batchnorm.d:140-154).  With a data-parallel world > 1 the real gradients are exchanged, the zero
    ones are not, and the projected running statistics are averaged over ranks so that all replicas stay identical."""
    def count(world):
        H.reset()
        H.seed(3)
        H.set_data_parallel_world(world)
        x, y = H.float32((4, 3, 8, 8)), H.float32((4, 5))
        out = H.data_source(x).conv2d(8, (3, 3), padding=(1, 1), use_bias=False).batch_norm().relu().dense(5).softmax()
        net = H.Network([x], [out])
        loss = H.cross_entropy(out.train_output, y) + net.param_loss
        upd = H.Updater(H.SGD, [loss], network=net, hyper=[H.float32((), [0.1]), H.float32((), [0.9])])
        plan_ops, _ = upd.plan_outputs()
        nodes = H.export(plan_ops)
        by_id = dict((n["id"], n) for n in nodes)
        reduced = [by_id[n["deps"][0]] for n in nodes if n["type"] == "allreduce"]
        return len(net.params), reduced
    n_params, reduced = count(2)
    assert n_params == 1 + 4 + 2                                   # conv w | scale, bias, mean, var | dense w, b
    # 5 real gradients (conv w, scale, bias, dense w, dense b) + the 2 running statistics; never a bare zero variable
    assert len(reduced) == 5 + 2
    assert all(r["type"] not in ("variable", "constant") for r in reduced)
    assert count(1)[1] == []


def test_no_backend_means_no_evaluation():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    x = H.float32((2,), [1, 2])
    with pytest.raises(H.HostError) as e:
        H.evaluate([x + x])
    assert "no backend" in str(e.value) or "no CPU" in str(e.value)
