"""Multi-GPU host logic: one process per GPU (torchrun), torch.distributed only as plumbing.

* Sampling shards independent images across ranks with NO collective: rank r
  takes a contiguous range; the device generator is keyed by the GLOBAL image index, so the
  union of the shards is bit-identical to a single-GPU run.
* Training is data parallel: every rank owns a libddpm handle; the NCCL communicator used for the
  gradient / BatchNorm-statistics all-reduces lives inside libddpm and is bootstrapped here by
  broadcasting its 128-byte unique id over the existing process group (gloo or nccl).
"""
from __future__ import annotations

import os
from typing import Tuple

import numpy as np


def env_rank_world() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (1-process defaults)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of ``n_total`` items owned by ``rank``; sizes differ by at most 1,
    earlier ranks take the remainder."""
    if world < 1 or not (0 <= rank < world) or n_total < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(n_total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def broadcast_bytes(payload: bytes, nbytes: int, src: int = 0) -> bytes:
    """Broadcast a fixed-size byte string from ``src`` over the default process group."""
    import torch
    import torch.distributed as td

    if not td.is_initialized() or td.get_world_size() == 1:
        return payload
    dev = "cuda" if td.get_backend() == "nccl" else "cpu"
    buf = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    if td.get_rank() == src:
        buf.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    td.broadcast(buf, src=src)
    return bytes(buf.cpu().numpy().tobytes())


def init_data_parallel(handle, sync_bn: bool = True) -> None:
    """Create libddpm's NCCL communicator across the ranks of the default process group."""
    import torch.distributed as td

    from . import capi

    if not td.is_initialized() or td.get_world_size() == 1:
        return
    rank, world = td.get_rank(), td.get_world_size()
    uid = capi.comm_unique_id() if rank == 0 else b"\x00" * 128
    uid = broadcast_bytes(uid, 128, src=0)
    handle.comm_init(uid, rank, world, sync_bn=sync_bn)


def gather_shards(local: np.ndarray, n_total: int) -> np.ndarray:
    """Host-side gather of per-rank sample shards to every rank (optional; the sampler itself
    needs no collective)."""
    import torch
    import torch.distributed as td

    if not td.is_initialized() or td.get_world_size() == 1:
        return local
    world = td.get_world_size()
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    maxn = max(e - b for b, e in sizes)
    dev = "cuda" if td.get_backend() == "nccl" else "cpu"
    pad = np.zeros((maxn,) + local.shape[1:], local.dtype)
    pad[:local.shape[0]] = local
    t = torch.from_numpy(pad).to(dev)
    outs = [torch.empty_like(t) for _ in range(world)]
    td.all_gather(outs, t)
    return np.concatenate([o.cpu().numpy()[:e - b] for o, (b, e) in zip(outs, sizes)], axis=0)
