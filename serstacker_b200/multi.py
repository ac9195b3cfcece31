"""Multi-GPU plumbing of the stacking path: frames shard across ranks (one process per GPU), every rank runs the
whole per-frame loop on its shard with no data-path collective, and the epilogue reduces the accumulator pair
(sum w*I, sum w) to one rank.  The reference is single-process ("frames shard, one reduce of
(sum wI, sum w)"); the running mean A = sum(w I) / sum(w) of c_weigthed_average (c_frame_accumulation.cc:20-129) is
associative in that sum form, so the result equals single-process stacking up to fp32 summation order.

On GPUs the reduce is the library's own (ssk_stack_reduce / ssk_acc_reduce in the C ABI: one ncclReduce group on the
pipeline's stream); torch.distributed only carries the 128-byte NCCL id between the ranks.  reduce_sum_form is the same
combine over torch.distributed (gloo in the CPU tests of the host logic)."""
import ctypes as C


def shard_frames(nframes, rank, world):
    """Contiguous shard [lo, hi) of `nframes` frames for `rank` of `world`; sizes differ by at most one."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("shard_frames: bad rank %d of %d" % (rank, world))
    base, extra = divmod(int(nframes), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def reduce_sum_form(acc_sum, wsum, nframes, dst=0, group=None):
    """Reduce (sum w*I, sum w) and the accumulated-frame count to rank `dst`, in place.
    acc_sum / wsum: torch tensors (CUDA for NCCL, CPU for gloo).  Returns the total frame count on `dst`
    (the local count elsewhere)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return int(nframes)
    n = torch.tensor([int(nframes)], dtype=torch.int64, device=acc_sum.device)
    dist.reduce(acc_sum, dst, op=dist.ReduceOp.SUM, group=group)
    dist.reduce(wsum, dst, op=dist.ReduceOp.SUM, group=group)
    dist.reduce(n, dst, op=dist.ReduceOp.SUM, group=group)
    return int(n.item())


class _DeviceView:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes // 4,), "typestr": "<f4", "data": (ptr, False), "version": 3}


class NcclComm:
    """An ncclComm_t owned through the C ABI (ssk_nccl_comm_create): what a C++ host passes to ssk_stack_reduce."""

    def __init__(self, unique_id, nranks, rank):
        from . import capi
        self._h = C.c_void_p()
        buf = C.create_string_buffer(bytes(unique_id), 128)
        capi.check(capi.lib.ssk_nccl_comm_create(buf, int(nranks), int(rank), C.byref(self._h)))

    @property
    def handle(self):
        return self._h

    def destroy(self):
        from . import capi
        if self._h:
            capi.check(capi.lib.ssk_nccl_comm_destroy(self._h))
            self._h = C.c_void_p()


def nccl_unique_id():
    """ncclGetUniqueId through the C ABI -> 128 bytes."""
    from . import capi
    buf = C.create_string_buffer(128)
    capi.check(capi.lib.ssk_nccl_get_unique_id(buf))
    return buf.raw


def create_comm_from_torch(group=None):
    """One NcclComm per rank of an initialised torch.distributed group: rank 0 draws the id, the group broadcasts it.
    The current CUDA device must be the rank's GPU."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    return NcclComm(box[0], world, rank)


def reduce_pipeline(pipe, comm, dst=0):
    """Epilogue of a sharded run through the C ABI: rank `dst`'s pipeline accumulator becomes the stack of all ranks'
    frames (ssk_stack_reduce).  Returns accumulated_frames() (the total on `dst`)."""
    from . import capi
    capi.check(capi.lib.ssk_stack_reduce(pipe._h, comm.handle, int(dst)))
    return pipe.accumulated_frames()


def combine_pipeline(pipe, device, dst=0, group=None):
    """The same epilogue over torch.distributed (zero-copy views of the accumulator's device buffers): kept as the
    cross-check of ssk_stack_reduce."""
    import torch
    import torch.distributed as dist
    from . import capi
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return pipe.accumulated_frames()
    acc_h = capi.lib.ssk_stack_accumulator(pipe._h)
    pa, pw, ba, bw = C.c_void_p(), C.c_void_p(), C.c_int64(), C.c_int64()
    capi.check(capi.lib.ssk_acc_device_state(acc_h, C.byref(pa), C.byref(pw), C.byref(ba), C.byref(bw)))
    pipe.sync()
    capi.check(capi.lib.ssk_acc_to_sum_form(acc_h))
    ta = torch.as_tensor(_DeviceView(pa.value, ba.value), device=device)
    tw = torch.as_tensor(_DeviceView(pw.value, bw.value), device=device)
    torch.cuda.synchronize(device)
    total = reduce_sum_form(ta, tw, pipe.accumulated_frames(), dst, group)
    torch.cuda.synchronize(device)
    capi.check(capi.lib.ssk_acc_from_sum_form(acc_h, total))
    return total
