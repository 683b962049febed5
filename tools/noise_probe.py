"""Diagnostics: run the same one-step WRN-16-4 training plan twice per configuration and print how much the parameter updates of
the two runs differ (median / worst tensor, relative to the update).  Environment knobs are read when a plan is built, so every
configuration runs in this one process."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import dopt_b200 as db  # noqa: E402
from dopt_b200 import host as H  # noqa: E402
import test_plan_gpu as T  # noqa: E402

FUSE, GRAPH, INTERIOR = db._lib.PLAN_FUSE, db._lib.PLAN_CUDA_GRAPH, db._lib.PLAN_BF16_INTERIOR


def run(flags, math, steps=1):
    H.reset()
    H.set_math(math)
    H.set_plan_flags(flags)
    loss, extra, net, feed = T._wrn(16, 4, 32, 16, 10)()
    upd = H.Updater(H.SGD, [loss] + extra, network=net, hyper=[H.float32((), [0.05]), H.float32((), [0.9])])
    init = [p.get().copy() for p in net.params]
    outs = [upd.step(feed(s)) for s in range(steps)]
    return float(outs[0][0]), [p.get().copy() for p in net.params], init


def diffs(pa, pb, init):
    out = []
    for a, b, i0 in zip(pa, pb, init):
        du = float(np.linalg.norm((a - i0).astype(np.float64)))
        if du > 1e-12:
            out.append(float(np.linalg.norm((a - b).astype(np.float64))) / du)
    return np.array(out)


def main():
    assert H.init()
    configs = [
        ("bf16 interior (default)", {}, FUSE | GRAPH | INTERIOR, db.MATH_BF16),
        ("bf16 interior, PDL off", {"DOPT_B200_PDL": "0"}, FUSE | GRAPH | INTERIOR, db.MATH_BF16),
        ("bf16 interior, no side stream", {"DOPT_B200_NO_SIDE_STREAM": "1"}, FUSE | GRAPH | INTERIOR, db.MATH_BF16),
        ("bf16 interior, no producer stats", {"DOPT_B200_NO_PRODUCER_STATS": "1"}, FUSE | GRAPH | INTERIOR, db.MATH_BF16),
        ("bf16 interior, no halo", {"DOPT_B200_HALO": "0", "DOPT_B200_WG_HALO": "0"}, FUSE | GRAPH | INTERIOR, db.MATH_BF16),
        ("bf16 operands, fp32 storage", {}, FUSE | GRAPH, db.MATH_BF16),
        ("fp32", {}, FUSE | GRAPH, db.MATH_FP32),
    ]
    if os.environ.get("NOISE_PROBE_LOSS"):
        # the step-0 loss of repeated runs (forward only) as a cheap discriminator between modes
        knobs = ["DOPT_B200_NO_WG_ARENA", "DOPT_B200_NO_SIDE_STREAM", "DOPT_B200_PDL", "DOPT_B200_NO_PACK_MERGE", "DOPT_B200_NO_FILTER_STAGE",
                 "DOPT_B200_HALO", "DOPT_B200_NO_MSUM", "DOPT_B200_CTAS_PER_SM", "DOPT_B200_NO_DEFER_FINISH"]
        for name, env in [("base", {}), ("no wg arena", {"DOPT_B200_NO_WG_ARENA": "1"}), ("no side stream", {"DOPT_B200_NO_SIDE_STREAM": "1"}),
                          ("pdl off", {"DOPT_B200_PDL": "0"}), ("no side, pdl off", {"DOPT_B200_NO_SIDE_STREAM": "1", "DOPT_B200_PDL": "0"}),
                          ("no halo", {"DOPT_B200_HALO": "0"}), ("no msum", {"DOPT_B200_NO_MSUM": "1"}),
                          ("one cta/sm", {"DOPT_B200_CTAS_PER_SM": "1"}), ("no defer finish", {"DOPT_B200_NO_DEFER_FINISH": "1"})]:
            for k in knobs:
                os.environ.pop(k, None)
            os.environ["DOPT_B200_NO_PRODUCER_STATS"] = "1"
            os.environ.update(env)
            losses = [run(FUSE | GRAPH | INTERIOR, db.MATH_BF16)[0] for _ in range(8)]
            print("%-18s %s" % (name, " ".join("%.6f" % l for l in losses)), flush=True)
        return
    if os.environ.get("NOISE_PROBE_DETAIL"):
        # where along the net do two runs of the same plan part ways?  (parameters in creation order: input -> output)
        os.environ["DOPT_B200_NO_PRODUCER_STATS"] = "1"
        res = [run(FUSE | GRAPH | INTERIOR, db.MATH_BF16) for _ in range(4)]
        base = res[0]
        other = None
        for r in res[1:]:
            if np.median(diffs(base[1], r[1], base[2])) > 1e-3:
                other = r
        if other is None:
            print("all four runs agree")
        else:
            for k, (a, b, i0) in enumerate(zip(base[1], other[1], base[2])):
                du = float(np.linalg.norm((a - i0).astype(np.float64)))
                print(k, a.shape, "%.3g" % (float(np.linalg.norm((a - b).astype(np.float64))) / max(du, 1e-30)), "update norm %.3g" % du)
        os.environ.pop("DOPT_B200_NO_PRODUCER_STATS")
        configs = [
            ("no producer stats, no add absorb", {"DOPT_B200_NO_PRODUCER_STATS": "1", "DOPT_B200_NO_ADD_ABSORB": "1"}, FUSE | GRAPH | INTERIOR, db.MATH_BF16),
            ("no producer stats, no add stage", {"DOPT_B200_NO_PRODUCER_STATS": "1", "DOPT_B200_NO_ADD_STAGE": "1"}, FUSE | GRAPH | INTERIOR, db.MATH_BF16),
            ("no producer stats, no defer finish", {"DOPT_B200_NO_PRODUCER_STATS": "1", "DOPT_B200_NO_DEFER_FINISH": "1"}, FUSE | GRAPH | INTERIOR, db.MATH_BF16),
            ("no producer stats, no wg arena", {"DOPT_B200_NO_PRODUCER_STATS": "1", "DOPT_B200_NO_WG_ARENA": "1"}, FUSE | GRAPH | INTERIOR, db.MATH_BF16),
            ("no producer stats, no filter stage", {"DOPT_B200_NO_PRODUCER_STATS": "1", "DOPT_B200_NO_FILTER_STAGE": "1"}, FUSE | GRAPH | INTERIOR, db.MATH_BF16),
            ("no producer stats, fused only (no graph)", {"DOPT_B200_NO_PRODUCER_STATS": "1"}, FUSE | INTERIOR, db.MATH_BF16),
        ]
    for name, env, flags, math in configs:
        for k in ("DOPT_B200_PDL", "DOPT_B200_NO_SIDE_STREAM", "DOPT_B200_NO_PRODUCER_STATS", "DOPT_B200_HALO", "DOPT_B200_WG_HALO",
                  "DOPT_B200_NO_ADD_ABSORB", "DOPT_B200_NO_ADD_STAGE", "DOPT_B200_NO_DEFER_FINISH", "DOPT_B200_NO_WG_ARENA",
                  "DOPT_B200_NO_FILTER_STAGE"):
            os.environ.pop(k, None)
        os.environ.update(env)
        res = [run(flags, math) for _ in range(4)]
        d = [diffs(res[0][1], res[i][1], res[0][2]) for i in (1, 2, 3)]
        print("%-36s loss %s   median %s   worst %s" % (name, ["%.6f" % r[0] for r in res], ["%.2g" % np.median(x) for x in d],
                                                      ["%.2g" % x.max() for x in d]), flush=True)


if __name__ == "__main__":
    main()
