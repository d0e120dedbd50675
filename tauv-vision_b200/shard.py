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


def gather_frames(local, n_frames: int, dst: int = 0, group=None):
    """Gather per-frame result tensors onto ``dst`` in frame order with ONE collective and ONE device->host copy.

    ``local`` maps names to tensors whose leading dimension is this rank's frames (``frame_range(rank, world,
    n_frames)``); trailing shapes and dtypes are the same on every rank.  Every rank packs its tensors into one byte
    buffer (padded to the longest block), ``torch.distributed.gather`` moves the buffers — over NVLink when the tensors
    are on the GPUs (NCCL), through gloo for CPU tensors — and ``dst`` copies the stacked buffers to the host once and
    cuts them into numpy arrays.  Returns ``{name: array[n_frames, ...]}`` on ``dst`` and None elsewhere; without a
    process group, the local tensors as numpy arrays.  (``gather_host`` pickles every array through
    ``gather_object``: 4 ms per step at 2 GPUs for 1.7 MB per rank, against 0.3 ms this way.)"""
    import torch
    import torch.distributed as dist
    keys = sorted(local)
    if not (dist.is_available() and dist.is_initialized()):
        return {k: local[k].detach().cpu().numpy() for k in keys}
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = frame_range(rank, world, n_frames)
    blocks = [frame_range(r, world, n_frames) for r in range(world)]
    max_frames = max(h - l for l, h in blocks)
    device = local[keys[0]].device
    meta, off = [], 0
    for k in keys:
        v = local[k]
        if v.shape[0] != hi - lo:
            raise ValueError(f"'{k}' has {v.shape[0]} frames, this rank owns {hi - lo}")
        per = int(np.prod(v.shape[1:], dtype=np.int64)) * v.element_size()
        meta.append((k, tuple(v.shape[1:]), v.dtype, per, off))
        off += (max_frames * per + 15) // 16 * 16
    buf = torch.zeros((max(off, 16),), dtype=torch.uint8, device=device)
    for (k, _, _, per, o) in meta:
        n = (hi - lo) * per
        if n:
            buf[o:o + n] = local[k].detach().contiguous().reshape(-1).view(torch.uint8)
    bucket = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, bucket, dst=dst, group=group)
    if rank != dst:
        return None
    host = torch.stack(bucket).cpu()  # one device->host copy of world x bytes
    out = {}
    for (k, shape, dtype, per, o) in meta:
        parts = []
        for r, (l, h) in enumerate(blocks):
            n = (h - l) * per
            parts.append(host[r, o:o + n].view(dtype).reshape((h - l,) + shape))
        out[k] = torch.cat(parts, dim=0).numpy()
    return out
