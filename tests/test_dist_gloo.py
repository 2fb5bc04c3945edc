"""N > 1 host logic on CPU: world_size-2 gloo.  Rank 0 holds the whole batch, scatters
spectra + side info by stream, each rank runs its shard (the CPU emulation of the kernel
schedule stands in for the GPU), rank 0 gathers the PCM and checks it against the oracle
run on the unsharded batch."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, S, T, C, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import aacjs_b200  # noqa: F401  (registers the package)
    from aacjs_b200 import sharding
    from oracle import oracle as O
    from tests import emul
    from tools import workloads as W

    lo, hi = sharding.stream_range(S, world, rank)
    n = hi - lo
    full_spec = full_info = full_pcm = None
    if rank == 0:
        w = W.make(5, S, T, C, seed=7, shape_prev_mode="carried")
        full_spec = torch.from_numpy(w["spectra"])
        full_info = torch.from_numpy(w["info"].view(np.uint8).reshape(S, T, C, 8).copy())
        full_pcm = torch.empty((S, T, 1024, C))
    spec = torch.empty((n, T, C, 1024))
    info = torch.empty((n, T, C, 8), dtype=torch.uint8)
    sharding.scatter_streams(full_spec, spec, S)
    sharding.scatter_streams(full_info, info, S)
    ov = np.zeros((n, C, 1024), np.float32)
    pcm = emul.process(spec.numpy(), info.numpy().view(W.INFO_DTYPE).reshape(n, T, C), None, None, ov, 4, 0, 5)
    sharding.gather_streams(torch.from_numpy(pcm), full_pcm, S)
    if rank == 0:
        ref, _ = O.process(w["spectra"], w["info"], sample_index=4)
        q.put(float(np.abs(full_pcm.numpy().astype(np.float64) - ref).max()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("S", [5, 2])
def test_scatter_compute_gather_world2(S):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300 + S
    procs = [ctx.Process(target=_worker, args=(r, 2, port, S, 7, 2, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) <= 1e-5


def test_stream_ranges_partition_the_batch():
    sys.path.insert(0, ROOT)
    import aacjs_b200  # noqa: F401
    from aacjs_b200 import sharding

    for S in (1, 2, 7, 8, 256, 1000):
        for world in (1, 2, 4, 8):
            ranges = [sharding.stream_range(S, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == S
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1
