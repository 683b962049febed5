"""bench.py's host-side arithmetic (no GPU): the algorithmic bytes behind `roofline_classes` come from the exported dopt
graph and SURVEY.md section 8(d)'s per-element figures."""
import numpy as np

import bench
from dopt_b200 import host as H


def test_class_rooflines_from_the_exported_graph():
    H.init()
    x, y, net, upd = bench.build_wrn(H, 4, depth=10, width=2)
    outs, _ = upd.plan_outputs()
    nodes = H.export(outs)
    loss_id = [n["id"] for n in nodes if n["op"].h == outs[0].h][0]
    by_id = dict((n["id"], n) for n in nodes)
    bns = [n for n in nodes if n["type"] == "batchNormTrain"]
    v_bn = sum(int(np.prod(by_id[n["deps"][0]]["shape"])) for n in bns)
    n_params = sum(p.volume for p in net.params)
    # WRN-10-2: one block per group -> 3 residual adds, 2 batch norms per block + the final one
    assert len(bns) == 7
    us = {"batchNormTrain": 10.0, "batchNormGrad": 20.0, "add": 5.0, "update": 2.0}
    n = {"batchNormTrain": 7, "batchNormGrad": 7, "add": 3, "update": 4}
    r = bench.class_rooflines(nodes, loss_id, us, n, n_params, 6500.0)
    assert r["batchNormTrain"]["alg_bytes_per_step"] == 2 * v_bn * 4
    assert r["batchNormGrad"]["alg_bytes_per_step"] == 3 * v_bn * 4
    assert r["optimiser"]["alg_bytes_per_step"] == 5 * n_params * 4
    # residual sums: N x (16k) x 32x32, N x (32k) x 16x16, N x (64k) x 8x8 with k = 2
    v_add = 4 * (32 * 32 * 32 + 64 * 16 * 16 + 128 * 8 * 8)
    assert r["residual_add"]["alg_bytes_per_step"] == 3 * v_add * 4
    for k in r:
        assert abs(r[k]["frac"] - r[k]["achieved"] / 6500.0) < 1e-12
        assert abs(r[k]["achieved"] - r[k]["alg_bytes_per_step"] / (r[k]["us_per_step"] * 1e-6) / 1e9) < 1e-6
    # a profile whose `add` launches are not exactly the forward residual sums reports no residual_add class
    n["add"] = 5
    assert "residual_add" not in bench.class_rooflines(nodes, loss_id, us, n, n_params, 6500.0)
    H.reset()


def test_both_arms_name_the_same_workload():
    assert "WRN-28-10" in bench.workload_name() and "batch 128/GPU" in bench.workload_name()
