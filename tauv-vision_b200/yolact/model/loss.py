"""YOLACT anchor matching and regression targets — the target-encode half of
``tauv_vision.yolact.model.loss.loss`` (/root/reference/src/tauv_vision/yolact/model/loss.py:16-22 and
:62-66).  The loss arithmetic that consumes these (cross-entropy with hard negatives, smooth-L1, mask BCE)
needs autograd and is out of scope.  Kernel: csrc/yolact_boxes.cu (match_anchors_kernel).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from ... import _lib


@dataclass
class AnchorMatch:
    match_index: torch.Tensor   # [B,N] i64 — best truth per prior (first max on ties)
    match_iou: torch.Tensor     # [B,N] f32
    positive_match: torch.Tensor  # [B,N] bool  (iou >= config.iou_pos_threshold)
    negative_match: torch.Tensor  # [B,N] bool  (iou <= config.iou_neg_threshold)
    box_target: torch.Tensor    # [B,N,4] f32 — box_encode(truth_box[match_index], anchor) where positive, 0 elsewhere


def match_anchors(anchor: torch.Tensor, truth_box: torch.Tensor, truth_valid: torch.Tensor, config) -> AnchorMatch:
    """iou_matrix(anchor, truth_box) * valid -> max over truths -> thresholds -> box_encode, one pass,
    without the [B,N,M] temporaries (loss.py:16-22, :62-66)."""
    dev = _lib.require_cuda(anchor, truth_box, truth_valid)
    anc, tb = _lib.f32c(anchor), _lib.f32c(truth_box)
    if anc.dim() != 3 or anc.shape[0] != 1 or anc.shape[2] != 4:
        raise ValueError(f"anchor must be [1,N,4]; got {tuple(anc.shape)}")
    B, M = tb.shape[:2]
    N = anc.shape[1]
    tv = truth_valid.contiguous()
    tv = tv.view(torch.uint8) if tv.dtype == torch.bool else tv.to(torch.uint8)
    mi = torch.empty((B, N), dtype=torch.int64, device=dev)
    miou = torch.empty((B, N), dtype=torch.float32, device=dev)
    pos = torch.empty((B, N), dtype=torch.bool, device=dev)
    neg = torch.empty((B, N), dtype=torch.bool, device=dev)
    tgt = torch.empty((B, N, 4), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().tauv_yolact_match_anchors(
            _lib.fptr(anc), _lib.fptr(tb), _lib.u8ptr(tv), B, N, M, float(config.iou_pos_threshold),
            float(config.iou_neg_threshold), float(config.box_variances[0]), float(config.box_variances[1]),
            _lib.i64ptr(mi), _lib.fptr(miou), _lib.u8ptr(pos.view(torch.uint8)), _lib.u8ptr(neg.view(torch.uint8)),
            _lib.fptr(tgt), _lib.stream_ptr(dev)))
    return AnchorMatch(mi, miou, pos, neg, tgt)
