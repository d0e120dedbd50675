"""YOLACT box arithmetic — drop-in for ``tauv_vision.yolact.model.boxes``
(/root/reference/src/tauv_vision/yolact/model/boxes.py): ``box_xy_swap`` (:6-12), ``box_to_corners``
(:15-27), ``corners_to_box`` (:30-42), ``box_encode`` (:45-52), ``box_decode`` (:55-61), ``iou_matrix``
(:64-85), ``box_to_mask`` (:88-103).  Kernels: csrc/yolact_boxes.cu.

``config`` is duck-typed: only ``config.box_variances`` is read, as in the reference.
"""
from __future__ import annotations

import torch

from ... import _lib


def _boxes(t: torch.Tensor) -> torch.Tensor:
    if t.dim() != 3 or t.shape[-1] != 4:
        raise ValueError(f"boxes must be [n_batch, n, 4]; got {tuple(t.shape)}")
    return _lib.f32c(t)


def box_xy_swap(box: torch.Tensor) -> torch.Tensor:
    """(y,x,h,w) <-> (x,y,w,h); pure layout, any device: the reference's data loader, collate and plots call it
    on CPU tensors (datasets/segmentation_dataset/segmentation_dataset.py:119, yolact/scripts/train.py:143-145,
    utils/plot.py:49,65) — there is no kernel behind it, so nothing here insists on CUDA."""
    return box[:, :, [1, 0, 3, 2]]


def box_to_corners(box: torch.Tensor) -> torch.Tensor:
    """(y,x,h,w) -> (min_y,min_x,max_y,max_x).  Layout helper off the hot path (the kernels form corners
    in registers); kept as a thin torch expression with the reference's operation order, on any device."""
    half = box[:, :, 2:] / 2
    return torch.cat((box[:, :, :2] - half, box[:, :, :2] + half), dim=-1)


def corners_to_box(corners: torch.Tensor) -> torch.Tensor:
    """(min_y,min_x,max_y,max_x) -> (y,x,h,w); layout helper, any device (boxes.py:30-42)."""
    return torch.cat(((corners[:, :, :2] + corners[:, :, 2:]) / 2, corners[:, :, 2:] - corners[:, :, :2]), dim=-1)


def _codec(fn_name: str, x: torch.Tensor, anchor: torch.Tensor, config) -> torch.Tensor:
    dev = _lib.require_cuda(x, anchor)
    x, anchor = _boxes(x), _boxes(anchor)
    B = max(x.shape[0], anchor.shape[0])
    if x.shape[0] != B:
        x = x.expand(B, -1, -1).contiguous()
    N = x.shape[1]
    if anchor.shape[1] != N or anchor.shape[0] not in (1, B):
        raise RuntimeError(f"anchor {tuple(anchor.shape)} does not broadcast against {tuple(x.shape)}")
    out = torch.empty((B, N, 4), dtype=torch.float32, device=dev)
    v0, v1 = float(config.box_variances[0]), float(config.box_variances[1])
    with torch.cuda.device(dev):
        _lib.check(getattr(_lib.load(), fn_name)(_lib.fptr(x), _lib.fptr(anchor), B, N, anchor.shape[0], v0, v1,
                                                 _lib.fptr(out), _lib.stream_ptr(dev)))
    return out


def box_encode(box: torch.Tensor, anchor: torch.Tensor, config) -> torch.Tensor:
    """g_yx = (b_yx - a_yx)/(v0*a_hw), g_hw = log(b_hw/a_hw)/v1   — boxes.py:45-52."""
    return _codec("tauv_yolact_box_encode", box, anchor, config)


def box_decode(box_encoding: torch.Tensor, anchor: torch.Tensor, config) -> torch.Tensor:
    """yx = a_yx + e_yx*v0*a_hw, hw = a_hw*exp(e_hw*v1)   — boxes.py:55-61."""
    return _codec("tauv_yolact_box_decode", box_encoding, anchor, config)


def iou_matrix(box_a: torch.Tensor, box_b: torch.Tensor) -> torch.Tensor:
    """[Ba,Na,4] x [Bb,Nb,4] -> [B,Na,Nb] with broadcasting batch dims   — boxes.py:64-85."""
    dev = _lib.require_cuda(box_a, box_b)
    a, b = _boxes(box_a), _boxes(box_b)
    Ba, Bb = a.shape[0], b.shape[0]
    if not (Ba == Bb or Ba == 1 or Bb == 1):
        raise RuntimeError(f"batch dims {Ba} and {Bb} do not broadcast")
    B = max(Ba, Bb)
    out = torch.empty((B, a.shape[1], b.shape[1]), dtype=torch.float32, device=dev)
    if out.numel() == 0:
        return out
    with torch.cuda.device(dev):
        _lib.check(_lib.load().tauv_iou_matrix(_lib.fptr(a), _lib.fptr(b), Ba, Bb, a.shape[1], b.shape[1],
                                               _lib.fptr(out), _lib.stream_ptr(dev)))
    return out


def box_to_mask(box: torch.Tensor, img_size) -> torch.Tensor:
    """[4] (y,x,h,w) -> [H,W] {0,1}, inclusive crop on integer pixel coordinates   — boxes.py:88-103."""
    dev = _lib.require_cuda(box)
    bx = _lib.f32c(box).reshape(4)
    H, W = int(img_size[0]), int(img_size[1])
    out = torch.empty((H, W), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().tauv_box_to_mask(_lib.fptr(bx), H, W, _lib.fptr(out), _lib.stream_ptr(dev)))
    return out
