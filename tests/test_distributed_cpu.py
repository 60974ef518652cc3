"""Host-side logic of the N > 1 path on CPU: world_size-2 gloo process group.

Covers what the host does for the slab decomposition -- slab arithmetic, the all-gather of the
64-byte cudaIpc handles (ordered by rank) and the sums of the per-rank reductions (TrueVariance,
FmaxPDF) that the reference performs with MPI_Reduce/MPI_Bcast.  The data path itself (peer
stores inside the FFT passes) is covered by the emulated-rank tests in tests/test_emulator.py
and by tests/test_gpu_multi.py on real GPUs.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pinocchio_b200.distributed import IPC_HANDLE_BYTES, allgather_bytes, allreduce_sum, slab_bounds


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        handle = bytes([rank + 1]) * IPC_HANDLE_BYTES
        got = allgather_bytes(handle)
        ok = len(got) == world and all(g == bytes([r + 1]) * IPC_HANDLE_BYTES for r, g in enumerate(got))
        tv = allreduce_sum(np.array([1.0 + rank, 10.0 * (rank + 1)]))
        ok &= bool(np.allclose(tv, [sum(1.0 + r for r in range(world)), sum(10.0 * (r + 1) for r in range(world))]))
        pdf = allreduce_sum(np.arange(210, dtype=np.int64) * (rank + 1))
        ok &= bool(np.array_equal(pdf, np.arange(210) * sum(r + 1 for r in range(world))))
        lo, hi = slab_bounds(64, rank, world)
        ok &= (hi - lo) == 64 // world and lo == rank * (64 // world)
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_handle_exchange_and_reductions():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def test_slab_bounds():
    assert slab_bounds(2048, 3, 8) == (768, 1024)
    assert slab_bounds(1024, 0, 1) == (0, 1024)
    with pytest.raises(ValueError):
        slab_bounds(100, 0, 3)
