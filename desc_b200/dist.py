"""Host-side multi-GPU plumbing: one process per GPU (torchrun), torch.distributed for the
rendezvous only.  The data-path collectives (all-reduce of the 2m partner-sum accumulators,
all-gather of S_vec shards and of the GCW node blocks) run inside libdesc_b200.so over NCCL.
"""
from __future__ import annotations

import os

import numpy as np


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def broadcast_bytes(payload, nbytes, src=0, device="cpu"):
    """Broadcast ``nbytes`` bytes from rank ``src`` to every rank of the default process group
    (any backend).  ``payload`` is only read on ``src``."""
    import torch
    import torch.distributed as dist
    t = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    if dist.get_rank() == src:
        t = torch.frombuffer(bytearray(payload), dtype=torch.uint8).clone().to(device)
        if t.numel() != nbytes:
            raise ValueError("payload has %d bytes, expected %d" % (t.numel(), nbytes))
    dist.broadcast(t, src)
    return bytes(t.cpu().numpy().tobytes())


def exchange_nccl_id(make_id, device="cpu"):
    """Rank 0 creates the 128-byte ncclUniqueId (``make_id()``), everybody gets it."""
    import torch.distributed as dist
    payload = make_id() if dist.get_rank() == 0 else None
    return broadcast_bytes(payload, 128, 0, device)


def shard_bounds(rowptr_all, world):
    """Slot-balanced contiguous edge ranges, same rule as k_shard_bounds in csrc/build.cu:
    boundary r is the first edge whose row pointer is >= r*m_cycle/world."""
    rowptr_all = np.asarray(rowptr_all, dtype=np.int64)
    m = rowptr_all.size - 1
    total = int(rowptr_all[-1])
    b = [0]
    for r in range(1, world):
        b.append(int(np.searchsorted(rowptr_all[:m], (total * r) // world, side="left")))
    b.append(m)
    return np.array(b, dtype=np.int64)
