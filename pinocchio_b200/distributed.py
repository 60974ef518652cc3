"""Process-group plumbing for the slab decomposition: one process per GPU.

The reference moves data between ranks with MPI; here the heavy exchange (the distributed-FFT
transposes) is done by the CUDA kernels themselves through peer memory, and the host only has to
(a) move one 64-byte cudaIpc handle per rank between the processes and (b) sum a few numbers
(Sum(delta^2) per radius, the 210-bin Fmax histogram).  Both go through ``torch.distributed``
(NCCL or gloo backend alike -- the payloads are tiny host tensors).
"""
from __future__ import annotations

import ctypes

import numpy as np

IPC_HANDLE_BYTES = 64


def slab_bounds(n: int, rank: int, nranks: int) -> tuple[int, int]:
    """[start, stop) of the slab owned by ``rank``: real space is split along x and k space
    along y, both in equal contiguous slabs (src/initialization.c:1317-1325)."""
    if nranks < 1 or n % nranks:
        raise ValueError("the grid side must be divisible by the number of ranks")
    lx = n // nranks
    return rank * lx, (rank + 1) * lx


def _dist():
    import torch.distributed as dist
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised (launch with torchrun or init_process_group)")
    return dist


def _device_for(group):
    import torch
    dist = _dist()
    backend = dist.get_backend(group)
    return torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")


def allgather_bytes(payload: bytes, group=None) -> list[bytes]:
    """All-gather equally sized byte strings, ordered by rank."""
    import torch
    dist = _dist()
    dev = _device_for(group)
    mine = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(dev)
    outs = [torch.empty_like(mine) for _ in range(dist.get_world_size(group))]
    dist.all_gather(outs, mine, group=group)
    return [bytes(o.cpu().numpy().tobytes()) for o in outs]


def allreduce_sum(a: np.ndarray, group=None) -> np.ndarray:
    """Sum a small host array over the ranks (MPI_Reduce + MPI_Bcast in the reference)."""
    import torch
    dist = _dist()
    dev = _device_for(group)
    t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


def connect_peers(pin) -> None:
    """Exchange the arena handles of all ranks and map the peers (pinb200_ipc_handle/connect)."""
    buf = (ctypes.c_ubyte * IPC_HANDLE_BYTES)()
    pin._ck(pin.lib.pinb200_ipc_handle(pin.h, buf))
    handles = allgather_bytes(bytes(buf), pin.group)
    if len(handles) != pin.nranks:
        raise RuntimeError(f"process group has {len(handles)} ranks, expected {pin.nranks}")
    blob = b"".join(handles)
    cbuf = (ctypes.c_ubyte * len(blob)).from_buffer_copy(blob)
    pin._ck(pin.lib.pinb200_connect(pin.h, cbuf))
    _dist().barrier(group=pin.group)
