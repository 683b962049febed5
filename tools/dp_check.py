"""Run under torchrun (N ranks, one GPU each): data-parallel SGD steps through the real NCCL gradient buckets, checked
against the single-process CPU oracle on the concatenated batch.  Prints PASS/FAIL on rank 0."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LOCAL_RANK = int(os.environ.get("LOCAL_RANK", "0"))
WORLD = int(os.environ.get("WORLD_SIZE", "1"))
RANK = int(os.environ.get("RANK", "0"))
# DOPT_B200_SYMM=1: the gradient buckets live in peer-mapped memory and are reduced by the library's own multicast kernel
# (dopt_b200/symm.py); torch's symmetric-memory allocator tells peers apart by device index, so all GPUs stay visible then
SYMM = os.environ.get("DOPT_B200_SYMM", "0") == "1"
if SYMM:
    DEV = LOCAL_RANK
else:
    DEV = 0
    os.environ["CUDA_VISIBLE_DEVICES"] = str(LOCAL_RANK)
# the gradient all-reduces get a fixed, small number of CTAs and the tensor-core kernels leave those SMs free (comm.cu); NCCL
# reads the variable when the process creates its first communicator, which torch.distributed does below
os.environ.setdefault("NCCL_MAX_NCHANNELS", os.environ.get("DOPT_B200_COMM_CHANNELS", "16"))
os.environ.setdefault("NCCL_MIN_NCHANNELS", os.environ["NCCL_MAX_NCHANNELS"])

import ctypes as C  # noqa: E402
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import dopt_b200 as db  # noqa: E402
from dopt_b200 import host as H  # noqa: E402
from oracle import graph_eval as G  # noqa: E402

F = np.float32
PER_RANK = 4


def build(batch):
    H.reset()
    H.seed(31)
    x = H.float32((batch, 16, 8, 8))
    y = H.float32((batch, 10))
    l = H.data_source(x).conv2d(32, (3, 3), padding=(1, 1), weight_decay=1e-3, use_bias=False).relu() \
        .conv2d(32, (3, 3), padding=(1, 1), weight_decay=1e-3).relu().dense(10).softmax()
    net = H.Network([x], [l])
    loss = H.cross_entropy(l.train_output, y) + net.param_loss
    upd = H.Updater(H.SGD, [loss], network=net, hyper=[H.float32((), [0.05]), H.float32((), [0.9])])
    return x, y, net, upd


def main():
    torch.cuda.set_device(DEV)
    dist.init_process_group("nccl", device_id=torch.device("cuda", DEV))
    assert H.init(), H.init_error()
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if RANK == 0:
        buf = C.create_string_buffer(128)
        db.check(db.lib.dopt_b200_comm_unique_id(buf))
        uid.copy_(torch.tensor(list(buf.raw), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    H.init_data_parallel(RANK, WORLD, bytes(uid.cpu().tolist()))
    keep = None
    if SYMM:
        from dopt_b200 import symm
        keep = symm.attach(64 << 20, torch.device("cuda", DEV))
        assert keep is not None, "symmetric memory did not come up"
    H.set_math(db.MATH_FP32)
    rng = np.random.RandomState(9)
    total = PER_RANK * WORLD
    data = [(rng.randn(total, 16, 8, 8).astype(F), np.eye(10, dtype=F)[rng.randint(0, 10, total)]) for _ in range(4)]
    x, y, net, upd = build(PER_RANK)
    lo, hi = RANK * PER_RANK, (RANK + 1) * PER_RANK
    for s in range(4):
        upd.step({x: data[s][0][lo:hi], y: data[s][1][lo:hi]})
    mine = [p.get() for p in net.params]
    stats = upd.stats()
    # replicas identical?
    ok = True
    for p in mine:
        t = torch.from_numpy(p).cuda()
        ref = t.clone()
        dist.broadcast(ref, 0)
        ok = ok and bool(torch.equal(t, ref))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)

    # second phase: a WRN with batch norm on the bf16 tensor-core path with bf16 interior activations (the production
    # configuration), so that every plan pass (relu / staging / add absorbed into batch norm, flat batch norms, gradient
    # buckets, filter staging, deferred filter-gradient finishes, batched reductions) runs under data parallelism.  Batch
    # statistics are per rank, so there is no single-process oracle for this one; what must hold is that the replicas stay
    # bit-identical -- every rank applies the same averaged gradients AND the same averaged running statistics -- and that
    # the loss is finite and falls.
    H.reset()
    H.set_data_parallel_world(WORLD)
    H.set_math(db.MATH_BF16)
    H.seed(13)
    H.set_plan_flags(db._lib.PLAN_FUSE | db._lib.PLAN_CUDA_GRAPH | db._lib.PLAN_BF16_INTERIOR)
    xw, yw = H.float32((8, 3, 16, 16)), H.float32((8, 10))
    preds = H.wide_resnet(xw, 10, 4).dense(10).softmax()
    netw = H.Network([xw], [preds])
    lossw = H.cross_entropy(preds.train_output, yw) + netw.param_loss
    updw = H.Updater(H.SGD, [lossw], network=netw, hyper=[H.float32((), [0.05]), H.float32((), [0.9])])
    rw = np.random.RandomState(100 + RANK)
    batches = [((rw.rand(8, 3, 16, 16) * 2 - 1).astype(F), np.eye(10, dtype=F)[rw.randint(0, 10, 8)]) for _ in range(2)]
    losses = [float(updw.step({xw: batches[s % 2][0], yw: batches[s % 2][1]})[0]) for s in range(6)]
    okw = all(np.isfinite(l) for l in losses) and losses[-1] < losses[0]
    differing = []
    for k_, p_ in enumerate(netw.params):
        t = torch.from_numpy(p_.get()).cuda()
        ref = t.clone()
        dist.broadcast(ref, 0)
        if not bool(torch.equal(t, ref)):
            differing.append((k_, tuple(p_.shape)))
    # (the running mean / variance of batch norm are averaged over the ranks by dopt.online's exchange, so they are identical too)
    bad = differing
    if RANK == 1 and (bad or not okw):
        print("dp_check WRN phase: losses", losses, "differing trainable params", bad[:8], flush=True)
    okw = okw and not bad
    flagw = torch.tensor([1 if okw else 0], device="cuda")
    dist.all_reduce(flagw, op=dist.ReduceOp.MIN)
    flag = torch.minimum(flag, flagw)
    wrn_launches = updw.stats()["launches"]
    db.check(db.lib.dopt_b200_comm_check())   # NCCL asynchronous errors, timed-out flag barriers of the multicast all-reduce
    if RANK == 0:
        H.set_data_parallel_world(1)
        x1, y1, net1, upd1 = build(total)
        oracle = G.UpdaterOracle(upd1)
        for s in range(4):
            oracle.step({x1: data[s][0], y1: data[s][1]})
        worst = 0.0
        for p, got in zip(net1.params, mine):
            want = oracle.value_of(p)
            worst = max(worst, float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-6)))
        good = bool(flag.item()) and worst < 1e-4
        print("dp_check world=%d exchange=%s replicas_identical=%s (incl. WRN-10-4 bf16: %s, %d launches) "
              "max_rel_err_vs_single_process_oracle=%.3g launches=%d -> %s"
              % (WORLD, "multicast" if keep is not None else "nccl", bool(flag.item()), bool(flagw.item()), wrn_launches, worst,
                 stats["launches"], "PASS" if good else "FAIL"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
