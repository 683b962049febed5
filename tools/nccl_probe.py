"""Run under torchrun: times the library's own all-reduce entry point (dopt_b200_allreduce -> ncclAllReduce on the
communicator of comm.cu) for the message sizes of the gradient buckets, alone on the GPU.  One line per size on rank 0."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LOCAL_RANK = int(os.environ.get("LOCAL_RANK", "0"))
WORLD = int(os.environ.get("WORLD_SIZE", "1"))
RANK = int(os.environ.get("RANK", "0"))
os.environ["CUDA_VISIBLE_DEVICES"] = str(LOCAL_RANK)
if os.environ.get("DOPT_B200_COMM_CHANNELS", "16") != "0":
    os.environ.setdefault("NCCL_MAX_NCHANNELS", os.environ.get("DOPT_B200_COMM_CHANNELS", "16"))
    os.environ.setdefault("NCCL_MIN_NCHANNELS", os.environ["NCCL_MAX_NCHANNELS"])

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import dopt_b200 as db  # noqa: E402
from dopt_b200 import host as H  # noqa: E402


def main():
    torch.cuda.set_device(0)
    dist.init_process_group("gloo")
    assert H.init(), H.init_error()
    uid = torch.zeros(128, dtype=torch.uint8)
    if RANK == 0:
        buf = C.create_string_buffer(128)
        db.check(db.lib.dopt_b200_comm_unique_id(buf))
        uid.copy_(torch.tensor(list(buf.raw), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    H.init_data_parallel(RANK, WORLD, bytes(uid.tolist()))
    stream = torch.cuda.current_stream().cuda_stream
    out = []
    for mb in (1.5, 6.0, 36.0):
        n = int(mb * (1 << 20) / 4)
        x = torch.ones(n, device="cuda", dtype=torch.float32)
        for _ in range(5):
            db.check(db.lib.dopt_b200_allreduce(C.c_void_p(x.data_ptr()), n, C.c_float(1.0), C.c_void_p(stream)))
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            db.check(db.lib.dopt_b200_allreduce(C.c_void_p(x.data_ptr()), n, C.c_float(1.0), C.c_void_p(stream)))
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 20
        out.append("%.1fMB %.0fus %.0fGB/s" % (mb, us, n * 4 / us / 1e3))
    if RANK == 0:
        print(os.environ.get("PROBE_TAG", ""), "world", WORLD, " | ".join(out), flush=True)
    dist.barrier()
    db.lib.dopt_b200_comm_destroy()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
