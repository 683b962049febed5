"""Data-parallel host logic on CPU: world_size 2 over gloo.  Each rank builds the SAME dopt graph with the data-parallel
world set to 2 (so dopt.online wraps gradients in `allreduce`), evaluates its half of the batch with the CPU oracle, and
the gradient exchange is done with a gloo all-reduce (mean) in place of the NCCL kernel.  The updated parameters must
(a) be identical on both ranks and (b) equal a single-process step on the concatenated batch when the loss is a mean over
the batch -- the sharding rule of SURVEY.md section 8(e)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F = np.float32


def _build(H, batch):
    H.reset()
    H.seed(21)
    x = H.float32((batch, 4, 6, 6))
    y = H.float32((batch, 3))
    out = H.data_source(x).conv2d(4, (3, 3), padding=(1, 1), weight_decay=1e-3).relu().dense(3).softmax()
    net = H.Network([x], [out])
    loss = H.cross_entropy(out.train_output, y) + net.param_loss
    upd = H.Updater(H.SGD, [loss], network=net, hyper=[H.float32((), [0.1]), H.float32((), [0.9])])
    return x, y, net, upd


def _data():
    rng = np.random.RandomState(5)
    return rng.randn(8, 4, 6, 6).astype(F), np.eye(3, dtype=F)[rng.randint(0, 3, 8)]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dopt_b200 import host as H
    from oracle import graph_eval as G
    H.init()
    H.set_data_parallel_world(world)
    x, y, net, upd = _build(H, 8 // world)
    plan_ops, dests = upd.plan_outputs()
    nodes = H.export(plan_ops)
    assert sum(1 for n in nodes if n["type"] == "allreduce") == len(net.params)

    # oracle evaluator with the collective plugged in where the `allreduce` nodes sit
    import oracle.dopt_ref as R
    orig = R.evaluate_op

    def eval_with_collective(op_type, inputs, attrs=None, out_shape=None):
        return orig(op_type, inputs, attrs, out_shape)
    values = G.snapshot_values(nodes)
    fs, ls = _data()
    lo, hi = rank * (8 // world), (rank + 1) * (8 // world)
    args = {x.serial: fs[lo:hi], y.serial: ls[lo:hi]}
    env = {}
    for n in nodes:
        if n["type"] in ("variable", "constant"):
            env[n["id"]] = np.asarray(args.get(n["id"], values.get(n["id"])), dtype=n["dtype"]).reshape(n["shape"])
        elif n["type"] == "allreduce":
            t = torch.from_numpy(np.ascontiguousarray(env[n["deps"][0]]).copy())
            dist.all_reduce(t)                       # sum over ranks ...
            env[n["id"]] = (t.numpy() / world).astype(F)   # ... times 1/world: the mean (what the NCCL kernel does)
        else:
            env[n["id"]] = np.asarray(eval_with_collective(n["type"], [env[d] for d in n["deps"]], n["attrs"], n["shape"]),
                                      dtype=n["dtype"]).reshape(n["shape"])
    new_params = [env[o.serial] for o, d in zip(plan_ops, dests) if d is not None][:len(net.params)]
    q.put((rank, [p.copy() for p in new_params]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_data_parallel_step_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for a, b in zip(results[0], results[1]):
        np.testing.assert_array_equal(a, b)          # replicas stay identical
    # single process, whole batch, no allreduce
    sys.path.insert(0, ROOT)
    from dopt_b200 import host as H
    from oracle import graph_eval as G
    H.init()
    H.set_data_parallel_world(1)
    x, y, net, upd = _build(H, 8)
    oracle = G.UpdaterOracle(upd)
    fs, ls = _data()
    oracle.step({x: fs, y: ls})
    for p, got in zip(net.params, results[0]):
        np.testing.assert_allclose(got, oracle.value_of(p), rtol=2e-5, atol=2e-6)
    H.reset()


def test_rank_invariant_weight_decay_term_is_not_exchanged(monkeypatch):
    """dopt.online.exchange: the gradient of a regularised parameter is add(lossGradient, weightDecayTerm); the second operand
    depends only on the parameters, which every rank holds identically, so only the first is all-reduced and the term is added
    behind the collective (mean(a_r + b) = mean(a_r) + b).  DOPT_B200_NO_EXCHANGE_SPLIT=1 restores allreduce(add(..))."""
    sys.path.insert(0, ROOT)
    from dopt_b200 import host as H
    H.init()

    def operand_types(split):
        if split:
            monkeypatch.delenv("DOPT_B200_NO_EXCHANGE_SPLIT", raising=False)
        else:
            monkeypatch.setenv("DOPT_B200_NO_EXCHANGE_SPLIT", "1")
        H.set_data_parallel_world(2)
        x, y, net, upd = _build(H, 4)
        plan_ops, dests = upd.plan_outputs()
        nodes = H.export(plan_ops)
        by_id = {n["id"]: n for n in nodes}
        ops = sorted(by_id[n["deps"][0]]["type"] for n in nodes if n["type"] == "allreduce")
        n_params = len(net.params)
        H.set_data_parallel_world(1)
        H.reset()
        return ops, n_params
    plain, n_params = operand_types(False)
    split, _ = operand_types(True)
    assert len(plain) == len(split) == n_params
    assert "convolutionFiltersGrad" not in plain and "add" in plain       # conv filter and dense weight carry weight decay
    assert "convolutionFiltersGrad" in split
    assert split.count("add") < plain.count("add")


def _vote_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib.util
    # dopt_b200/symm.py without the package import (the vote itself needs neither the CUDA library nor a GPU)
    spec = importlib.util.spec_from_file_location("symm_only", os.path.join(ROOT, "dopt_b200", "symm.py"))
    symm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(symm)
    q.put((rank, symm.all_agree(True), symm.all_agree(rank != 1), symm.all_agree(False)))
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_choice_is_a_vote_over_all_ranks():
    """symm.attach: the multicast exchange is used only when the symmetric buffer came up on EVERY rank -- a rank that fell
    back to NCCL on its own would wait in ncclAllReduce for peers sitting in the multicast kernel's flag barrier."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31700 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_vote_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, all_ok, one_failed, none_ok in results:
        assert all_ok is True and one_failed is False and none_ok is False
