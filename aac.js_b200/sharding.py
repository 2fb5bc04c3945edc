"""Frame-batch sharding across the GPUs of one box.

Channel-frames only couple along time inside one (stream, channel) chain, so a batch
shards by *stream* with no data-path collective: rank r owns streams [lo_r, hi_r) and
its own overlap state.  When the batch starts out on one rank (the decoder host feeds
rank 0), the only communication is a scatter of spectra + side info and a gather of
PCM -- grouped point-to-point sends/receives (NCCL has no scatter/gather primitive;
`torch.distributed.batch_isend_irecv` maps to ncclGroupStart/ncclSend/ncclRecv/End).
No reduction of any kind exists on this path.  Backend-agnostic: NCCL on GPUs, gloo
in the CPU tests.

`share_from_root` is the fused alternative on NVLink/NVSwitch boxes: every rank maps rank 0's
buffers through CUDA IPC and hands the *peer* addresses of its stream range straight to
`aacfb_process_device`.  The synthesis kernel then is scatter, compute and gather in one: its
TMA bulk copies pull the spectrum rows (and `cp.async` the side info) from rank 0's HBM over
NVLink, and its 16-byte PCM stores land in rank 0's output buffer, row by row, overlapped with
the arithmetic -- no staging copy, no separate collective, both link directions busy at once.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def stream_range(n_streams: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced split of streams: the first (n % world) ranks get one extra."""
    base, extra = divmod(n_streams, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def scatter_streams(full: torch.Tensor | None, local: torch.Tensor, n_streams: int, root: int = 0) -> None:
    """Send rows [lo_r, hi_r) of `full` (dim 0 = streams, on root) into `local` on every rank."""
    rank, world = dist.get_rank(), dist.get_world_size()
    ops = []
    if rank == root:
        for r in range(world):
            lo, hi = stream_range(n_streams, world, r)
            if r == root:
                local.copy_(full[lo:hi])
            elif hi > lo:
                ops.append(dist.P2POp(dist.isend, full[lo:hi], r))
    else:
        lo, hi = stream_range(n_streams, world, rank)
        if hi > lo:
            ops.append(dist.P2POp(dist.irecv, local, root))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def gather_streams(local: torch.Tensor, full: torch.Tensor | None, n_streams: int, root: int = 0) -> None:
    """Inverse of scatter_streams: collect every rank's rows into `full` on root."""
    rank, world = dist.get_rank(), dist.get_world_size()
    ops = []
    if rank == root:
        for r in range(world):
            lo, hi = stream_range(n_streams, world, r)
            if r == root:
                full[lo:hi].copy_(local)
            elif hi > lo:
                ops.append(dist.P2POp(dist.irecv, full[lo:hi], r))
    else:
        lo, hi = stream_range(n_streams, world, rank)
        if hi > lo:
            ops.append(dist.P2POp(dist.isend, local, root))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


_ipc_open: dict = {}   # IPC handle bytes -> device pointer of the mapping in this process


def share_from_root(t: torch.Tensor | None, root: int = 0) -> int:
    """Every rank gets the ADDRESS of root's CUDA tensor `t` as seen from its own GPU: the
    allocation is exported with cudaIpcGetMemHandle on root and opened on every other rank with
    ITS device current (cudaIpcOpenMemHandle maps into the current device's address space and
    enables peer access, i.e. NVLink loads/stores from kernels running there).  On root this is
    `t.data_ptr()`.  One process per GPU on one box; `t` must stay alive on root while in use.
    (torch's own CUDA-IPC tensor sharing opens the handle with the EXPORTING device current, which
    makes the alias usable for copies but not for kernels running on another GPU.)"""
    from cuda.bindings import driver as drv
    from cuda.bindings import runtime as rt

    def ok(res):
        err, *out = res
        if int(err) != 0:
            raise RuntimeError(f"CUDA error {err}")
        return out[0] if len(out) == 1 else out

    rank = dist.get_rank()
    obj = [None]
    if rank == root:
        assert t.is_cuda and t.is_contiguous()
        base, _size = ok(drv.cuMemGetAddressRange(t.data_ptr()))
        handle = ok(rt.cudaIpcGetMemHandle(int(base)))
        obj[0] = (bytes(handle.reserved), t.data_ptr() - int(base))
    dist.broadcast_object_list(obj, src=root)
    if rank == root:
        return t.data_ptr()
    raw, offset = obj[0]
    if raw not in _ipc_open:
        ok(rt.cudaSetDevice(torch.cuda.current_device()))
        h = rt.cudaIpcMemHandle_t()
        h.reserved = raw
        _ipc_open[raw] = int(ok(rt.cudaIpcOpenMemHandle(h, rt.cudaIpcMemLazyEnablePeerAccess)))
    return _ipc_open[raw] + offset
