"""YOLACT prediction-head outputs in the layout their consumers read — the tail of
``tauv_vision.yolact.model.prediction_head.PredictionHead.forward``
(/root/reference/src/tauv_vision/yolact/model/prediction_head.py:111-113, :122-124, :137-140) and the concatenation over the
FPN levels of ``Yolact.forward`` (model.py:55-58).

The reference turns every level's NCHW convolution output into ``[B, H*W*A, C]`` with ``permute(0, 2, 3, 1).reshape(...)``
(a transposed copy per level), applies ``tanh`` to the mask coefficients (another pass) and ``torch.cat``s the levels (a
third copy).  ``pack_heads`` writes each level straight into its slice of the final tensor: one read and one write per
element, tanh on the way, with autograd (the backward is the inverse transposition).  Kernel: csrc/yolact_heads.cu.
The priors that go with these rows come from ``anchors.all_anchors`` (built once on the device and cached, instead of the
reference's CPU build + copy per forward, model.py:47-48).
"""
from __future__ import annotations

import ctypes
from typing import Sequence, Tuple

import torch

from ... import _lib


def _ptr_array(tensors):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


class _PackHeads(torch.autograd.Function):
    @staticmethod
    def forward(ctx, channels_per_prior, tanh_act, *levels):
        dev = levels[0].device
        B, CH = levels[0].shape[:2]
        hw = [int(t.shape[2] * t.shape[3]) for t in levels]
        hw_arr = (ctypes.c_int32 * len(hw))(*hw)
        out = torch.empty((B, sum(hw) * CH // channels_per_prior, channels_per_prior), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().tauv_yolact_pack_heads(_ptr_array(levels), hw_arr, len(levels), B, CH, int(tanh_act),
                                                          _lib.fptr(out), _lib.stream_ptr(dev)))
        ctx.shapes = [tuple(t.shape) for t in levels]
        ctx.tanh_act = bool(tanh_act)
        if tanh_act:
            ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, grad):
        dev = grad.device
        g = grad.to(torch.float32).contiguous()
        y = ctx.saved_tensors[0] if ctx.tanh_act else None
        grads = [torch.empty(s, dtype=torch.float32, device=dev) for s in ctx.shapes]
        B, CH = ctx.shapes[0][:2]
        hw_arr = (ctypes.c_int32 * len(grads))(*[s[2] * s[3] for s in ctx.shapes])
        with torch.cuda.device(dev):
            _lib.check(_lib.load().tauv_yolact_pack_heads_backward(
                _lib.fptr(g), _lib.fptr(y) if y is not None else None, hw_arr, len(grads), B, CH, int(ctx.tanh_act),
                _ptr_array(grads), _lib.stream_ptr(dev)))
        return (None, None, *grads)


def pack_head(levels: Sequence[torch.Tensor], channels_per_prior: int, tanh: bool = False) -> torch.Tensor:
    """``torch.cat([t.permute(0, 2, 3, 1).reshape(B, -1, C) for t in levels], dim=1)`` (and ``tanh`` of it) for the NCHW
    outputs ``[B, A*C, H_l, W_l]`` of one head over the FPN levels -> ``[B, sum_l H_l*W_l*A, C]``."""
    if not levels:
        raise ValueError("no levels")
    dev = _lib.require_cuda(*levels)
    lv = [_lib.f32c(t) for t in levels]
    B, CH = lv[0].shape[:2]
    if any(t.dim() != 4 or t.shape[0] != B or t.shape[1] != CH for t in lv) or CH % int(channels_per_prior) != 0:
        raise ValueError(f"levels must be [B, A*C, H, W] with the same B and A*C (C = {channels_per_prior}); "
                         f"got {[tuple(t.shape) for t in lv]}")
    del dev
    return _PackHeads.apply(int(channels_per_prior), bool(tanh), *lv)


def pack_heads(classification_levels: Sequence[torch.Tensor], box_encoding_levels: Sequence[torch.Tensor],
               mask_coeff_levels: Sequence[torch.Tensor], config) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """The three heads' per-level convolution outputs (``_classification_layer`` / ``_box_encoding_layer`` /
    ``_mask_coeff_layer`` of prediction_head.py:110, :121, :136, applied to every FPN level) ->
    ``(classification [B,N,n_classes+1], box_encoding [B,N,4], mask_coeff [B,N,n_prototype_masks])`` exactly as
    ``Yolact.forward`` returns them (model.py:55-60)."""
    return (pack_head(classification_levels, int(config.n_classes) + 1),
            pack_head(box_encoding_levels, 4),
            pack_head(mask_coeff_levels, int(config.n_prototype_masks), tanh=True))
