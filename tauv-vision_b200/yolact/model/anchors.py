"""YOLACT prior boxes — drop-in for ``tauv_vision.yolact.model.anchors.get_anchor``
(/root/reference/src/tauv_vision/yolact/model/anchors.py:9-41), generated on the device.

The reference rebuilds the anchors on the CPU and copies them to the GPU on every forward
(model.py:47-48); ``all_anchors`` builds all levels once, in place, and caches per configuration.
Ordering inside a level is aspect-major exactly as the reference produces it (it disagrees with the
prediction head's position-major flattening for more than one aspect ratio — reproduced, not fixed).
"""
from __future__ import annotations

import ctypes
from math import sqrt
from typing import Sequence, Tuple

import numpy as np
import torch

from ... import _lib


def _level_hw(fpn_i: int, config) -> Tuple[list, list]:
    # anchors.py:25-28 — Python doubles, then stored as fp32 by torch.full
    in_size = (config.in_h + config.in_w) / 2
    scale = config.anchor_scales[fpn_i]
    hs = [float(np.float32((scale / in_size) * sqrt(ar))) for ar in config.anchor_aspect_ratios]
    ws = [float(np.float32((scale / in_size) / sqrt(ar))) for ar in config.anchor_aspect_ratios]
    return hs, ws


def _build(levels: Sequence[Tuple[int, Tuple[int, int]]], config, device) -> torch.Tensor:
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if dev.type != "cuda":
        raise RuntimeError("tauv_vision_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
    A = len(config.anchor_aspect_ratios)
    hs = (ctypes.c_int * len(levels))(*[int(s[0]) for _, s in levels])
    ws = (ctypes.c_int * len(levels))(*[int(s[1]) for _, s in levels])
    hw = []
    for fpn_i, _ in levels:
        h, w = _level_hw(fpn_i, config)
        hw += h + w
    hw_arr = (ctypes.c_float * len(hw))(*hw)
    n = sum(A * int(s[0]) * int(s[1]) for _, s in levels)
    out = torch.empty((1, n, 4), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().tauv_yolact_anchors(hs, ws, len(levels), A, hw_arr, _lib.fptr(out),
                                                   _lib.stream_ptr(dev)))
    return out


def get_anchor(fpn_i: int, fpn_size, config, device=None) -> torch.Tensor:
    """[1, A*H*W, 4] (y,x,h,w) priors of FPN level ``fpn_i``   — anchors.py:9-41."""
    return _build([(fpn_i, (int(fpn_size[0]), int(fpn_size[1])))], config, device)


_CACHE: dict = {}


def all_anchors(fpn_sizes, config, device=None, cache: bool = True) -> torch.Tensor:
    """All levels concatenated, [1, sum_l A*H_l*W_l, 4]   — the torch.cat of model.py:47-58.

    The priors depend only on the level sizes and four configuration fields, so they are generated once per
    (sizes, configuration, device) and the same read-only tensor is handed out afterwards (the reference rebuilds
    them on the CPU and copies them to the GPU on every forward, model.py:47-48).  ``cache=False`` builds a fresh
    tensor (for callers that write into it)."""
    levels = [(i, (int(s[0]), int(s[1]))) for i, s in enumerate(fpn_sizes)]
    if not cache:
        return _build(levels, config, device)
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    key = (tuple(s for _, s in levels), tuple(float(v) for v in config.anchor_scales),
           tuple(float(v) for v in config.anchor_aspect_ratios), float(config.in_h), float(config.in_w), dev.type,
           dev.index if dev.index is not None else torch.cuda.current_device())
    hit = _CACHE.get(key)
    if hit is None:
        hit = _CACHE[key] = _build(levels, config, dev)
    return hit
