"""Frame sharding across the GPUs of one box.

Every function on the path is per frame (CenterNet top-k is per frame, decode.py:267-269; ``nms`` is
single-frame, nms.py:14-17; target encode and anchor matching are per frame), so N GPUs each take a
contiguous block of the batch and run the same kernels.  There is NO collective on the data path: the only
exchange is the final host-side gather of the (small) packed results, in frame order.  One process per GPU
(``torch.distributed``; NCCL on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np


def frame_range(rank: int, world: int, n_frames: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of the batch owned by ``rank``: blocks differ by at most one frame and the
    first ``n_frames % world`` ranks take the longer ones (256 frames on 8 GPUs -> 32 each)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    if n_frames < 0:
        raise ValueError("n_frames must be >= 0")
    base, extra = divmod(n_frames, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_frames(tensor, rank: int, world: int):
    """The rank's block of a batch-major tensor (a view, no copy)."""
    lo, hi = frame_range(rank, world, tensor.shape[0])
    return tensor[lo:hi]


def concat_host(parts: List[Optional[Dict[str, Optional[np.ndarray]]]]) -> Dict[str, Optional[np.ndarray]]:
    """Concatenate per-rank host dictionaries (``PackedDetections.to_host()`` / YOLACT keep lists) along the
    frame axis, in rank order = frame order.  Ranks that own no frames contribute None / empty arrays."""
    parts = [p for p in parts if p is not None]
    if not parts:
        return {}
    out: Dict[str, Optional[np.ndarray]] = {}
    for key in parts[0]:
        vals = [p[key] for p in parts]
        if any(v is None for v in vals):
            if not all(v is None for v in vals):
                raise ValueError(f"'{key}' is present on some ranks only")
            out[key] = None
        else:
            out[key] = np.concatenate(vals, axis=0)
    return out


def gather_host(local: Dict[str, Optional[np.ndarray]], dst: int = 0, group=None):
    """Host gather of the per-rank packed results onto ``dst`` (returns None elsewhere).  Single-process
    (no initialised process group) returns ``local`` unchanged."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    bucket = [None] * world if rank == dst else None
    dist.gather_object(local, bucket, dst=dst, group=group)
    if rank != dst:
        return None
    return concat_host(bucket)
