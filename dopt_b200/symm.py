"""Plumbing for dopt_b200_comm_set_symmetric (include/dopt_b200.h): obtains one peer-mapped buffer per rank plus the NVSwitch
multicast mapping over all of them from torch.distributed's symmetric-memory allocator (CUDA VMM + multicast objects underneath)
and hands the raw addresses to the library.  The gradient-bucket arenas of every plan created afterwards live in that buffer
and are reduced by the library's own multimem kernel (csrc/comm.cu) instead of ncclAllReduce.

torch's allocator identifies peers by CUDA device index, so all GPUs must be visible in every process and rank r works on
device LOCAL_RANK (with one visible device per process every rank would claim index 0 and the rendezvous refuses)."""
import ctypes as C


def attach(pool_bytes, device, group=None):
    """Returns an object that keeps the memory alive (drop it only after the last plan is gone and after detach()), or
    None when symmetric memory / multicast is not available here -- the library then keeps using NCCL."""
    import torch
    import torch.distributed as dist
    import torch.distributed._symmetric_memory as symm_mem

    from . import check, lib

    group = group or dist.group.WORLD
    try:
        n = (int(pool_bytes) + 1023) // 1024 * 256
        t = symm_mem.empty(n, dtype=torch.float32, device=device)
        hdl = symm_mem.rendezvous(t, group)
        off = int(getattr(hdl, "offset", 0) or 0)          # the tensor's offset inside the allocation the handle describes
        mc = int(hdl.multicast_ptr)
        ptrs = [int(p) + off for p in hdl.buffer_ptrs]
        pads = [int(p) for p in hdl.signal_pad_ptrs]
        if mc == 0 or len(pads) != hdl.world_size:
            return None
        mc += off
        if ptrs[hdl.rank] != t.data_ptr():
            raise RuntimeError("symmetric handle does not describe the tensor (%#x vs %#x)" % (ptrs[hdl.rank], t.data_ptr()))
        t.zero_()
        torch.cuda.synchronize()
        dist.barrier(group)
    except Exception as e:   # no NVSwitch multicast, no fd passing in this container, one visible device per process ...
        import sys
        sys.stderr.write("dopt_b200.symm: symmetric memory not available (%r); using NCCL\n" % (e,))
        return None
    arr = (C.c_void_p * len(pads))(*pads)
    check(lib.dopt_b200_comm_set_symmetric(C.c_void_p(ptrs[hdl.rank]), C.c_void_p(mc), C.c_size_t(n * 4), arr, len(pads),
                                           C.c_size_t(int(hdl.signal_pad_size))))
    return (t, hdl)


def detach():
    from . import lib
    lib.dopt_b200_comm_set_symmetric(None, None, 0, None, 0, 0)
