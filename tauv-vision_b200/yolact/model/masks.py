"""YOLACT mask assembly — drop-in for ``tauv_vision.yolact.model.masks.assemble_mask``
(/root/reference/src/tauv_vision/yolact/model/masks.py:8-21).  Kernels: csrc/yolact_mask.cu.

The prototype x coefficient contraction runs on the tensor cores (bf16 operands, fp32 accumulate), so the
pre-sigmoid logits carry bf16 input rounding: |error| <= 1e-2 absolute on the logits (north-star tolerance),
i.e. <= 2.5e-3 on the mask values.
"""
from __future__ import annotations

from typing import Optional

import torch

from ... import _lib


def assemble_mask(mask_prototype: torch.Tensor, mask_coeff: torch.Tensor, box: Optional[torch.Tensor],
                  return_logits: bool = False):
    """mask[i] = sigmoid(sum_p coeff[i,p]*proto[p]) * box_to_mask(box[i])   — masks.py:8-21.
    mask_prototype [P,H,W], mask_coeff [n,P], box [n,4] or None -> [n,H,W]."""
    dev = _lib.require_cuda(mask_prototype, mask_coeff, box)
    proto, coeff = _lib.f32c(mask_prototype), _lib.f32c(mask_coeff)
    if proto.dim() != 3 or coeff.dim() != 2 or coeff.shape[1] != proto.shape[0]:
        raise ValueError(f"mask_prototype {tuple(proto.shape)} / mask_coeff {tuple(coeff.shape)} must be [P,H,W] / [n,P]")
    P, H, W = proto.shape
    n = coeff.shape[0]
    bx = _lib.f32c(box) if box is not None else None
    if bx is not None and tuple(bx.shape) != (n, 4):
        raise ValueError(f"box must be [{n},4]; got {tuple(bx.shape)}")
    out = torch.empty((n, H, W), dtype=torch.float32, device=dev)
    logits = torch.empty((n, H, W), dtype=torch.float32, device=dev) if return_logits else None
    if n:
        with torch.cuda.device(dev):
            _lib.check(_lib.load().tauv_yolact_assemble_mask(_lib.fptr(proto), _lib.fptr(coeff), _lib.fptr(bx), n, P, H,
                                                             W, _lib.fptr(out), _lib.fptr(logits),
                                                             _lib.stream_ptr(dev)))
    return (out, logits) if return_logits else out


def assemble_mask_batched(mask_prototype: torch.Tensor, mask_coeff: torch.Tensor, detections, crop: bool = True,
                          out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """All frames at once from the fused detect() output: frame b gets n_keep[b] masks built from
    mask_coeff[b, keep[b,i]] and cropped to box[b,i].  mask_prototype [B,P,H,W], mask_coeff [B,N,P] ->
    [B,top_k,H,W]; rows >= n_keep[b] are left untouched.  No synchronisation."""
    dev = _lib.require_cuda(mask_prototype, mask_coeff, detections.keep)
    proto, coeff = _lib.f32c(mask_prototype), _lib.f32c(mask_coeff)
    B, P, H, W = proto.shape
    N = coeff.shape[1]
    top_k = detections.keep.shape[1]
    if out is None:
        out = torch.empty((B, top_k, H, W), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().tauv_yolact_assemble_mask_batched(
            _lib.fptr(proto), _lib.fptr(coeff), _lib.i64ptr(detections.keep), _lib.i32ptr(detections.n_keep),
            _lib.fptr(detections.box) if crop else None, B, N, P, H, W, top_k, _lib.fptr(out), _lib.stream_ptr(dev)))
    return out
