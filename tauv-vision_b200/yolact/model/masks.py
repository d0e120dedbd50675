"""YOLACT mask assembly — drop-in for ``tauv_vision.yolact.model.masks.assemble_mask``
(/root/reference/src/tauv_vision/yolact/model/masks.py:8-21).  Kernels: csrc/yolact_mask.cu.

The prototype x coefficient contraction runs on the tensor cores with every fp32 operand split into a bf16
(hi, lo) pair and three MMAs per product (fp32 accumulate), so the pre-sigmoid logits are within ~2e-5 of the
fp32 result — the tests assert <= 1e-4; the north star allows 1e-2 — and the mask values within 2.5e-3.
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


def _depth_u16(depth_mm: torch.Tensor, dev, dims: int) -> torch.Tensor:
    """The ROS mono16 depth image as 16-bit words on the device (uint16, or int16 holding the same bits)."""
    if depth_mm.dtype not in (torch.uint16, torch.int16):
        raise TypeError(f"depth_mm must be uint16 (mono16 millimetres; int16 with the same bits is accepted); got {depth_mm.dtype}")
    if depth_mm.dim() != dims:
        raise ValueError(f"depth_mm must have {dims} dimensions; got {tuple(depth_mm.shape)}")
    if depth_mm.device != dev:
        raise RuntimeError(f"depth_mm is on {depth_mm.device}, the prototypes on {dev} (there is no CPU path)")
    return depth_mm.contiguous()


def masked_depth_mean(mask_prototype: torch.Tensor, mask_coeff: torch.Tensor, box: Optional[torch.Tensor],
                      depth_mm: torch.Tensor, return_count: bool = False):
    """Mean camera depth under each detection's mask, fused with the mask assembly — what the ROS node computes with
    ``assemble_mask`` -> ``F.interpolate(mask[None], depth.shape)`` -> ``np.nanmean(np.where(mask > 0.5, depth, nan))``
    (/root/reference/src/tauv_vision/yolact/node/yolact_node.py:102-103,130-131,178), without writing a mask.
    mask_prototype [P,H,W], mask_coeff [n,P], box [n,4] or None, depth_mm [Hi,Wi] uint16 millimetres (0 = no reading)
    -> mean [n] float64 metres (NaN where the node would ``continue``), optionally the number of readings averaged."""
    dev = _lib.require_cuda(mask_prototype, mask_coeff, box)
    proto, coeff = _lib.f32c(mask_prototype), _lib.f32c(mask_coeff)
    if proto.dim() != 3 or coeff.dim() != 2 or coeff.shape[1] != proto.shape[0]:
        raise ValueError(f"mask_prototype {tuple(proto.shape)} / mask_coeff {tuple(coeff.shape)} must be [P,H,W] / [n,P]")
    P, H, W = proto.shape
    n = coeff.shape[0]
    bx = _lib.f32c(box) if box is not None else None
    if bx is not None and tuple(bx.shape) != (n, 4):
        raise ValueError(f"box must be [{n},4]; got {tuple(bx.shape)}")
    depth = _depth_u16(depth_mm, dev, 2)
    Hi, Wi = depth.shape
    mean = torch.full((n,), float("nan"), dtype=torch.float64, device=dev)
    count = torch.zeros((n,), dtype=torch.int64, device=dev)
    if n:
        lib = _lib.load()
        ws = torch.empty(lib.tauv_yolact_mask_depth_workspace_bytes(1, H, W, n), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.tauv_yolact_mask_depth(_lib.fptr(proto), _lib.fptr(coeff), _lib.fptr(bx), n, P, H, W,
                                                  depth.data_ptr(), Hi, Wi, _lib.dptr(mean), _lib.i64ptr(count),
                                                  ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev)))
    return (mean, count) if return_count else mean


def masked_depth_mean_batched(mask_prototype: torch.Tensor, mask_coeff: torch.Tensor, detections, depth_mm: torch.Tensor,
                              crop: bool = True, workspace: Optional[torch.Tensor] = None):
    """``masked_depth_mean`` for all frames at once from the fused detect() output.  mask_prototype [B,P,H,W],
    mask_coeff [B,N,P], depth_mm [B,Hi,Wi] uint16 -> (mean [B,top_k] float64 with NaN for rows >= n_keep[b] or
    without a reading, count [B,top_k] int64).  No synchronisation."""
    dev = _lib.require_cuda(mask_prototype, mask_coeff, detections.keep)
    proto, coeff = _lib.f32c(mask_prototype), _lib.f32c(mask_coeff)
    B, P, H, W = proto.shape
    N = coeff.shape[1]
    top_k = detections.keep.shape[1]
    depth = _depth_u16(depth_mm, dev, 3)
    if depth.shape[0] != B:
        raise ValueError(f"depth_mm must hold one image per frame: {tuple(depth.shape)} vs B={B}")
    Hi, Wi = depth.shape[1:]
    lib = _lib.load()
    need = lib.tauv_yolact_mask_depth_workspace_bytes(B, H, W, top_k)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=dev)
    mean = torch.empty((B, top_k), dtype=torch.float64, device=dev)
    count = torch.empty((B, top_k), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.tauv_yolact_mask_depth_batched(
            _lib.fptr(proto), _lib.fptr(coeff), _lib.i64ptr(detections.keep), _lib.i32ptr(detections.n_keep),
            _lib.fptr(detections.box) if crop else None, B, N, P, H, W, top_k, depth.data_ptr(), Hi, Wi,
            _lib.dptr(mean), _lib.i64ptr(count), workspace.data_ptr(), workspace.numel(), _lib.stream_ptr(dev)))
    return mean, count


_RESIZE_MODES = {"nearest": 0, "bilinear": 1}


def _resize_args(size, mode):
    if mode not in _RESIZE_MODES:
        raise ValueError(f"mode must be 'nearest' or 'bilinear'; got {mode!r}")
    out_h, out_w = (int(s) for s in size)
    if out_h <= 0 or out_w <= 0:
        raise ValueError(f"size must be positive; got {size}")
    return out_h, out_w, _RESIZE_MODES[mode]


def assemble_mask_binary(mask_prototype: torch.Tensor, mask_coeff: torch.Tensor, box: Optional[torch.Tensor], size,
                         mode: str = "nearest") -> torch.Tensor:
    """``F.interpolate(assemble_mask(proto, coeff, box)[None], size, mode=mode)[0] > 0.5`` as uint8 — the ROS node's
    mask (mode "nearest": /root/reference/src/tauv_vision/yolact/node/yolact_node.py:135 and its ``mask_np > 0.5`` at
    :178) and the evaluation script's (mode "bilinear": yolact/scripts/evaluate_batch.py:101-102), written once, one
    byte per pixel.  mask_prototype [P,H,W], mask_coeff [n,P], box [n,4] or None -> [n,size[0],size[1]] uint8."""
    dev = _lib.require_cuda(mask_prototype, mask_coeff, box)
    proto, coeff = _lib.f32c(mask_prototype), _lib.f32c(mask_coeff)
    if proto.dim() != 3 or coeff.dim() != 2 or coeff.shape[1] != proto.shape[0]:
        raise ValueError(f"mask_prototype {tuple(proto.shape)} / mask_coeff {tuple(coeff.shape)} must be [P,H,W] / [n,P]")
    out_h, out_w, m = _resize_args(size, mode)
    P, H, W = proto.shape
    n = coeff.shape[0]
    bx = _lib.f32c(box) if box is not None else None
    if bx is not None and tuple(bx.shape) != (n, 4):
        raise ValueError(f"box must be [{n},4]; got {tuple(bx.shape)}")
    out = torch.empty((n, out_h, out_w), dtype=torch.uint8, device=dev)
    if n:
        lib = _lib.load()
        ws = torch.empty(lib.tauv_yolact_mask_binary_workspace_bytes(1, H, W, n), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.tauv_yolact_mask_binary(_lib.fptr(proto), _lib.fptr(coeff), _lib.fptr(bx), n, P, H, W, out_h,
                                                   out_w, m, _lib.u8ptr(out), ws.data_ptr(), ws.numel(),
                                                   _lib.stream_ptr(dev)))
    return out


def assemble_mask_binary_batched(mask_prototype: torch.Tensor, mask_coeff: torch.Tensor, detections, size,
                                 mode: str = "nearest", crop: bool = True, out: Optional[torch.Tensor] = None,
                                 workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``assemble_mask_binary`` for all frames at once from the fused detect() output.  mask_prototype [B,P,H,W],
    mask_coeff [B,N,P] -> [B,top_k,size[0],size[1]] uint8; rows >= n_keep[b] are left untouched.  No synchronisation."""
    dev = _lib.require_cuda(mask_prototype, mask_coeff, detections.keep)
    proto, coeff = _lib.f32c(mask_prototype), _lib.f32c(mask_coeff)
    B, P, H, W = proto.shape
    N = coeff.shape[1]
    top_k = detections.keep.shape[1]
    out_h, out_w, m = _resize_args(size, mode)
    lib = _lib.load()
    need = lib.tauv_yolact_mask_binary_workspace_bytes(B, H, W, top_k)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=dev)
    if out is None:
        out = torch.empty((B, top_k, out_h, out_w), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.tauv_yolact_mask_binary_batched(
            _lib.fptr(proto), _lib.fptr(coeff), _lib.i64ptr(detections.keep), _lib.i32ptr(detections.n_keep),
            _lib.fptr(detections.box) if crop else None, B, N, P, H, W, top_k, out_h, out_w, m, _lib.u8ptr(out),
            workspace.data_ptr(), workspace.numel(), _lib.stream_ptr(dev)))
    return out
