"""YOLACT Fast NMS — drop-in for ``tauv_vision.yolact.model.nms.nms``
(/root/reference/src/tauv_vision/yolact/model/nms.py:7-29) plus the batched / fused forms the reference
lacks.  Kernels: csrc/yolact_nms.cu.

Order of the returned prior indices: confidence descending, ties by prior index ascending (the
reference leaves ties to torch.sort).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from ... import _lib


def _check(classification: torch.Tensor, box: torch.Tensor):
    if classification.dim() != 3 or box.dim() != 3 or box.shape[-1] != 4 or box.shape[:2] != classification.shape[:2]:
        raise ValueError(f"classification {tuple(classification.shape)} / box {tuple(box.shape)} must be "
                         "[B,N,C+1] / [B,N,4]")


def nms_batched(classification: torch.Tensor, box: torch.Tensor, top_k: int, iou_threshold: float,
                confidence_threshold: float, n_frames: Optional[int] = None):
    """Fast NMS for the first ``n_frames`` frames (default all): (keep [n_frames, top_k] i64, n_keep [n_frames] i32).
    No synchronisation."""
    dev = _lib.require_cuda(classification, box)
    _check(classification, box)
    cls, bx = _lib.f32c(classification), _lib.f32c(box)
    B, N, C1 = cls.shape
    nf = B if n_frames is None else int(n_frames)
    top_k = int(top_k)
    keep = torch.empty((nf, top_k), dtype=torch.int64, device=dev)
    n_keep = torch.empty((nf,), dtype=torch.int32, device=dev)
    lib = _lib.load()
    with torch.cuda.device(dev):
        ws = _lib.workspace(dev, lib.tauv_yolact_nms_workspace_bytes(nf, N, C1, top_k))
        _lib.check(lib.tauv_yolact_fast_nms(_lib.fptr(cls), _lib.fptr(bx), B, N, C1, nf, top_k, float(iou_threshold),
                                            float(confidence_threshold), _lib.i64ptr(keep), _lib.i32ptr(n_keep),
                                            ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev)))
    return keep, n_keep


def nms(classification: torch.Tensor, box: torch.Tensor, top_k: int, iou_threshold: float,
        confidence_threshold: float) -> torch.Tensor:
    """Reference signature: kept prior indices of FRAME 0 ONLY (nms.py:14-17,25), LongTensor[n_keep].
    The variable-length result forces one host read of n_keep."""
    keep, n_keep = nms_batched(classification, box, top_k, iou_threshold, confidence_threshold, n_frames=1)
    return keep[0, : int(n_keep[0])]


@dataclass
class YolactDetections:
    """Device-resident output of the fused head post-process (nothing here has synchronised).
    keep [B,top_k] i64 · n_keep [B] i32 · box [B,top_k,4] · score [B,top_k] · class_id [B,top_k] i32;
    rows >= n_keep[b] are unspecified."""
    keep: torch.Tensor
    n_keep: torch.Tensor
    box: torch.Tensor
    score: torch.Tensor
    class_id: torch.Tensor


def detect(classification: torch.Tensor, box_encoding: torch.Tensor, anchor: torch.Tensor, config, top_k: int,
           iou_threshold: float, confidence_threshold: float) -> YolactDetections:
    """box_decode + nms + class argmax of yolact_node.py:127-130,139-140 for every frame, decoding only the
    top_k priors.  ``config.box_variances`` is the only attribute read."""
    dev = _lib.require_cuda(classification, box_encoding, anchor)
    _check(classification, box_encoding)
    cls, enc, anc = _lib.f32c(classification), _lib.f32c(box_encoding), _lib.f32c(anchor)
    B, N, C1 = cls.shape
    if anc.shape[0] not in (1, B) or anc.shape[1] != N:
        raise RuntimeError(f"anchor {tuple(anc.shape)} does not broadcast against [{B},{N},4]")
    top_k = int(top_k)
    keep = torch.empty((B, top_k), dtype=torch.int64, device=dev)
    n_keep = torch.empty((B,), dtype=torch.int32, device=dev)
    kbox = torch.empty((B, top_k, 4), dtype=torch.float32, device=dev)
    kscore = torch.empty((B, top_k), dtype=torch.float32, device=dev)
    kclass = torch.empty((B, top_k), dtype=torch.int32, device=dev)
    lib = _lib.load()
    with torch.cuda.device(dev):
        ws = _lib.workspace(dev, lib.tauv_yolact_nms_workspace_bytes(B, N, C1, top_k))
        _lib.check(lib.tauv_yolact_detect(
            _lib.fptr(cls), _lib.fptr(enc), _lib.fptr(anc), B, N, C1, anc.shape[0],
            float(config.box_variances[0]), float(config.box_variances[1]), top_k, float(iou_threshold),
            float(confidence_threshold), _lib.i64ptr(keep), _lib.i32ptr(n_keep), _lib.fptr(kbox), _lib.fptr(kscore),
            _lib.i32ptr(kclass), ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev)))
    return YolactDetections(keep, n_keep, kbox, kscore, kclass)


def max_foreground_confidence(classification: torch.Tensor, with_argmax: bool = False):
    """softmax(classification)[..., 1:].max(-1) without materialising the softmax   — nms.py:9-10."""
    dev = _lib.require_cuda(classification)
    cls = _lib.f32c(classification)
    B, N, C1 = cls.shape
    score = torch.empty((B, N), dtype=torch.float32, device=dev)
    arg = torch.empty((B, N), dtype=torch.int32, device=dev) if with_argmax else None
    with torch.cuda.device(dev):
        _lib.check(_lib.load().tauv_yolact_scores(_lib.fptr(cls), B, N, C1, _lib.fptr(score), _lib.i32ptr(arg),
                                                  _lib.stream_ptr(dev)))
    return (score, arg) if with_argmax else score
