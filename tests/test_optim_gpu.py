"""Fused multi-tensor optimiser updates against the reference's update graphs evaluated op by op in float32
(online/source/dopt/online/{sgd,adam,amsgrad}.d).  Because the kernels keep the reference's operation order and never
contract into FMAs, the bar is BIT-EXACT."""
import ctypes as C

import numpy as np
import pytest
import torch

import dopt_b200 as db
from dopt_b200 import _lib
from oracle import dopt_ref as R

pytestmark = pytest.mark.gpu
F = np.float32
SIZES = [1, 3, 16, 4097, 640 * 640 * 9, 100 * 640 + 1]


def _params(tensors):
    arr = (_lib.Param * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i].w = t["w"].data_ptr()
        arr[i].g = t["g"].data_ptr() if t.get("g") is not None else None
        arr[i].s0 = t["s0"].data_ptr()
        arr[i].s1 = t["s1"].data_ptr() if "s1" in t else None
        arr[i].s2 = t["s2"].data_ptr() if "s2" in t else None
        arr[i].n = t["w"].numel()
    return arr


def _scalar(v):
    return torch.tensor([v], dtype=torch.float32, device="cuda")


@pytest.mark.parametrize("nesterov", [0, 1])
def test_sgd_bit_exact(nesterov):
    rng = np.random.RandomState(30)
    host = [{k: rng.randn(n).astype(F) for k in ("w", "g", "s0")} for n in SIZES]
    dev = [{k: torch.from_numpy(v).cuda() for k, v in t.items()} for t in host]
    lr, mu = _scalar(0.1), _scalar(0.9)
    for step in range(3):
        db.check(db.lib.dopt_b200_sgd_update(_params(dev), len(dev), lr.data_ptr(), mu.data_ptr(), nesterov, 1.0, None))
        for t in host:
            t["w"], t["s0"] = R.sgd_step(t["w"], t["g"], t["s0"], 0.1, 0.9, bool(nesterov))
    torch.cuda.synchronize()
    for h, d in zip(host, dev):
        np.testing.assert_array_equal(d["w"].cpu().numpy(), h["w"])
        np.testing.assert_array_equal(d["s0"].cpu().numpy(), h["s0"])


def test_sgd_zero_gradient_tensor_and_grad_scale():
    rng = np.random.RandomState(31)
    w, m, g = (rng.randn(1000).astype(F) for _ in range(3))
    dev = [{"w": torch.from_numpy(w).cuda(), "g": None, "s0": torch.from_numpy(m).cuda()},
           {"w": torch.from_numpy(w).cuda(), "g": torch.from_numpy(g).cuda(), "s0": torch.from_numpy(m).cuda()}]
    lr, mu = _scalar(0.05), _scalar(0.5)
    db.check(db.lib.dopt_b200_sgd_update(_params(dev), 2, lr.data_ptr(), mu.data_ptr(), 0, 0.25, None))
    torch.cuda.synchronize()
    nw, nm = R.sgd_step(w, np.zeros_like(w), m, 0.05, 0.5)       # BN running stats: gradient is a zero variable
    np.testing.assert_array_equal(dev[0]["w"].cpu().numpy(), nw)
    nw, nm = R.sgd_step(w, (g * F(0.25)).astype(F), m, 0.05, 0.5)   # data-parallel mean over 4 ranks
    np.testing.assert_array_equal(dev[1]["w"].cpu().numpy(), nw)
    np.testing.assert_array_equal(dev[1]["s0"].cpu().numpy(), nm)


@pytest.mark.parametrize("amsgrad", [0, 1])
def test_adam_amsgrad_bit_exact(amsgrad):
    rng = np.random.RandomState(32)
    host = []
    for n in SIZES:
        t = {"w": rng.randn(n).astype(F), "g": rng.randn(n).astype(F), "s0": rng.randn(n).astype(F) * F(0.1),
             "s1": np.abs(rng.randn(n)).astype(F) * F(0.01)}
        if amsgrad:
            t["s2"] = np.zeros(n, F)
        host.append(t)
    dev = [{k: torch.from_numpy(v).cuda() for k, v in t.items()} for t in host]
    alpha, beta1, beta2, eps = _scalar(1e-3), _scalar(0.9), _scalar(0.999), _scalar(1e-8)
    b1, b2 = _scalar(1.0), _scalar(1.0)
    hb1, hb2 = F(1.0), F(1.0)
    for step in range(3):
        db.check(db.lib.dopt_b200_adam_update(_params(dev), len(dev), alpha.data_ptr(), beta1.data_ptr(),
                                              beta2.data_ptr(), eps.data_ptr(), b1.data_ptr(), b2.data_ptr(), amsgrad,
                                              1.0, None))
        nb1 = nb2 = None
        for t in host:
            out = R.adam_step(t["w"], t["g"], t["s0"], t["s1"], hb1, hb2, 1e-3, 0.9, 0.999, 1e-8,
                              vhat=t.get("s2") if amsgrad else None)
            t["w"], t["s0"], t["s1"], nb1, nb2 = out[:5]
            if amsgrad:
                t["s2"] = out[5]
        hb1, hb2 = nb1, nb2
    torch.cuda.synchronize()
    assert b1.item() == hb1 and b2.item() == hb2
    for h, d in zip(host, dev):
        for k in h:
            if k == "g":
                continue
            got = d[k].cpu().numpy()
            if k == "w":
                # sqrt / div are IEEE-exact on both sides, so the whole chain is bit-exact
                np.testing.assert_array_equal(got, h[k])
            else:
                np.testing.assert_array_equal(got, h[k])
