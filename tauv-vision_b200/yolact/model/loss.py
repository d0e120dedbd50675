"""YOLACT loss — ``tauv_vision.yolact.model.loss.loss`` (/root/reference/src/tauv_vision/yolact/model/loss.py).

* anchor matching and regression targets (loss.py:16-22, :62-66): ``match_anchors`` — csrc/yolact_boxes.cu;
* classification term with hard-negative mining and the box term (loss.py:26-73), forward and backward, fused with
  the target-class lookup: ``class_box_loss`` — csrc/yolact_loss.cu (the class logits are read once per direction, the
  reference's per-frame Python loop of ~12 ATen kernels is gone);
* mask term (loss.py:75-121), forward and backward: ``mask_loss`` — csrc/yolact_loss.cu (no mask, resized truth mask or
  [SH,SW] temporary is materialised; the reference loops over the positives in Python);
* ``loss(prediction, truth, config)`` — the reference's entry point, same arguments and return value.
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


class _ClassBoxLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cls, enc, target, positive, negative, match_index, truth_cls, ratio):
        dev = cls.device
        B, N, C1 = cls.shape
        M = truth_cls.shape[1]
        lib = _lib.load()
        selected = torch.empty((B, N), dtype=torch.uint8, device=dev)
        pos_list = torch.empty((B, N), dtype=torch.int32, device=dev)
        sums = torch.empty((B, 2), dtype=torch.float64, device=dev)
        n_pos = torch.empty((B,), dtype=torch.int64, device=dev)
        ws = _lib.workspace(dev, lib.tauv_yolact_class_box_loss_workspace_bytes(B, N))
        with torch.cuda.device(dev):
            _lib.check(lib.tauv_yolact_class_box_loss(
                _lib.fptr(cls), _lib.fptr(enc), _lib.fptr(target), _lib.u8ptr(positive), _lib.u8ptr(negative),
                _lib.i64ptr(match_index), _lib.i64ptr(truth_cls), B, N, C1, M, int(ratio), _lib.u8ptr(selected),
                _lib.i32ptr(pos_list), _lib.dptr(sums), _lib.i64ptr(n_pos), ws.data_ptr(), ws.numel(),
                _lib.stream_ptr(dev)))
        losses = torch.empty((3,), dtype=torch.float32, device=dev)
        P = torch.empty((1,), dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):   # loss.py:54-57, :70-73: the normalisation, one small launch
            _lib.check(lib.tauv_yolact_loss_reduce(_lib.dptr(sums), _lib.i64ptr(n_pos), B, int(ratio), None, 0,
                                                   _lib.fptr(losses), _lib.i64ptr(P), _lib.stream_ptr(dev)))
        cls_loss, box_loss = losses[0], losses[1]
        ctx.save_for_backward(cls, enc, target, positive, selected, match_index, truth_cls, P)
        ctx.ratio = int(ratio)
        ctx.mark_non_differentiable(selected, pos_list, n_pos)
        return cls_loss, box_loss, selected, pos_list, n_pos

    @staticmethod
    def backward(ctx, g_cls, g_box, *_unused):
        cls, enc, target, positive, selected, match_index, truth_cls, P = ctx.saved_tensors
        dev = cls.device
        B, N, C1 = cls.shape
        need_cls, need_enc = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        grad_cls = torch.empty_like(cls) if need_cls else None
        grad_enc = torch.empty_like(enc) if need_enc else None
        gc = g_cls.to(torch.float32).contiguous().reshape(1) if need_cls else None
        gb = g_box.to(torch.float32).contiguous().reshape(1) if need_enc else None
        if need_cls or need_enc:
            with torch.cuda.device(dev):
                _lib.check(_lib.load().tauv_yolact_class_box_loss_backward(
                    _lib.fptr(cls), _lib.fptr(enc), _lib.fptr(target), _lib.u8ptr(positive), _lib.u8ptr(selected),
                    _lib.i64ptr(match_index), _lib.i64ptr(truth_cls), B, N, C1, truth_cls.shape[1], ctx.ratio,
                    _lib.i64ptr(P), _lib.fptr(gc) if need_cls else None, _lib.fptr(gb) if need_enc else None,
                    _lib.fptr(grad_cls) if need_cls else None, _lib.fptr(grad_enc) if need_enc else None,
                    _lib.stream_ptr(dev)))
        return grad_cls, grad_enc, None, None, None, None, None, None


@dataclass
class ClassBoxLoss:
    classification_loss: torch.Tensor  # fp32 scalar (loss.py:54-57)
    box_loss: torch.Tensor             # fp32 scalar (loss.py:70-73)
    selected: torch.Tensor             # [B,N] bool — positives and mined negatives
    pos_list: torch.Tensor             # [B,N] i32 — each frame's positives in prior order (first n_pos[b] entries)
    n_pos: torch.Tensor                # [B] i64


def class_box_loss(classification: torch.Tensor, box_encoding: torch.Tensor, match: AnchorMatch,
                   truth_classification: torch.Tensor, config) -> ClassBoxLoss:
    """Classification term with hard-negative mining (loss.py:26-56) and box term (loss.py:58-73) of the YOLACT loss
    from the class logits [B,N,C1], the predicted box encodings [B,N,4] and the anchor match; differentiable in
    ``classification`` and ``box_encoding``.  Where the reference's ``torch.topk`` leaves the choice among equal
    background confidences open, the lower prior index is taken."""
    dev = _lib.require_cuda(classification, box_encoding, truth_classification, match.positive_match)
    cls, enc = _lib.f32c(classification), _lib.f32c(box_encoding)
    if cls.dim() != 3 or enc.shape != cls.shape[:2] + (4,):
        raise ValueError(f"classification must be [B,N,C1] and box_encoding [B,N,4]; got {tuple(cls.shape)}, {tuple(enc.shape)}")
    tc = truth_classification.contiguous().to(torch.int64)
    pos = match.positive_match.contiguous().view(torch.uint8)
    neg = match.negative_match.contiguous().view(torch.uint8)
    cl, bl, sel, pl, n_pos = _ClassBoxLoss.apply(cls, enc, match.box_target, pos, neg, match.match_index, tc,
                                                 int(config.negative_example_ratio))
    return ClassBoxLoss(cl, bl, sel.view(torch.bool), pl, n_pos)


class _MaskLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coeff, proto, pos_list, n_pos, match_index, truth_box, seg, img_valid):
        dev = coeff.device
        B, N, K = coeff.shape
        PH, PW = proto.shape[-2:]
        SH, SW = seg.shape[-2:]
        lib = _lib.load()
        tsum = torch.empty((B, truth_box.shape[1]), dtype=torch.float64, device=dev)
        partial = torch.empty((B, lib.tauv_yolact_mask_loss_partials()), dtype=torch.float64, device=dev)
        recs = torch.empty((lib.tauv_yolact_mask_loss_records_bytes(B, N),), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.tauv_yolact_mask_loss(
                _lib.fptr(coeff), _lib.fptr(proto), _lib.i32ptr(pos_list), _lib.i64ptr(n_pos), _lib.i64ptr(match_index),
                _lib.fptr(truth_box), seg.data_ptr(), seg.element_size(), _lib.u8ptr(img_valid), B, N, K, truth_box.shape[1], PH, PW,
                SH, SW,
                _lib.dptr(tsum), recs.data_ptr(), _lib.dptr(partial), _lib.stream_ptr(dev)))
        losses = torch.empty((3,), dtype=torch.float32, device=dev)
        P = torch.empty((1,), dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):   # loss.py:117-120: the normalisation, one small launch
            _lib.check(lib.tauv_yolact_loss_reduce(None, _lib.i64ptr(n_pos), B, 0, _lib.dptr(partial), partial.numel(),
                                                   _lib.fptr(losses), _lib.i64ptr(P), _lib.stream_ptr(dev)))
        out = losses[2]
        ctx.save_for_backward(coeff, proto, pos_list, n_pos, match_index, truth_box, seg, img_valid, tsum, recs, P)
        return out

    @staticmethod
    def backward(ctx, g):
        coeff, proto, pos_list, n_pos, match_index, truth_box, seg, img_valid, tsum, recs, P = ctx.saved_tensors
        dev = coeff.device
        B, N, K = coeff.shape
        PH, PW = proto.shape[-2:]
        SH, SW = seg.shape[-2:]
        need_c, need_p = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gc = torch.empty_like(coeff) if need_c else None
        gp = torch.empty_like(proto) if need_p else None
        go = g.to(torch.float32).contiguous().reshape(1)
        if need_c or need_p:
            with torch.cuda.device(dev):
                _lib.check(_lib.load().tauv_yolact_mask_loss_backward(
                    _lib.fptr(coeff), _lib.fptr(proto), _lib.i32ptr(pos_list), _lib.i64ptr(n_pos),
                    _lib.i64ptr(match_index), _lib.fptr(truth_box), seg.data_ptr(), seg.element_size(), _lib.u8ptr(img_valid), B, N, K,
                    truth_box.shape[1], PH, PW, SH, SW, _lib.dptr(tsum), recs.data_ptr(), _lib.i64ptr(P), _lib.fptr(go),
                    _lib.fptr(gc) if need_c else None, _lib.fptr(gp) if need_p else None, _lib.stream_ptr(dev)))
        return gc, gp, None, None, None, None, None, None


def mask_loss(mask_coeff: torch.Tensor, mask_prototype: torch.Tensor, match: AnchorMatch, pos_list: torch.Tensor,
              n_pos: torch.Tensor, truth_box: torch.Tensor, truth_seg_map: torch.Tensor,
              truth_img_valid: torch.Tensor) -> torch.Tensor:
    """Mask term of the YOLACT loss (loss.py:75-121): per positive prior, BCE between the assembled mask and the
    bilinearly resized mask of its matched truth, inside the truth box and the valid image region, over the resized
    truth mask's area.  ``pos_list`` / ``n_pos`` are ``class_box_loss``'s list of positives.  ``mask_coeff`` [B,N,K],
    ``mask_prototype`` [B,K,PH,PW] (loss.py:82 sums over dim 0 of ``mask_prototype[b]``); differentiable in both, no
    mask is materialised.  Kernels: csrc/yolact_loss.cu (ymask_*)."""
    _lib.require_cuda(mask_coeff, mask_prototype, truth_box, truth_seg_map, truth_img_valid)
    coeff, proto = _lib.f32c(mask_coeff), _lib.f32c(mask_prototype)
    if coeff.dim() != 3 or proto.dim() != 4 or proto.shape[:2] != (coeff.shape[0], coeff.shape[2]):
        raise ValueError(f"mask_coeff must be [B,N,K] and mask_prototype [B,K,PH,PW]; got {tuple(coeff.shape)}, {tuple(proto.shape)}")
    seg = truth_seg_map.contiguous()   # read in place: uint8 (the reference's dataset), int32 or int64
    if seg.dtype not in (torch.uint8, torch.int32, torch.int64):
        seg = seg.to(torch.int32)
    valid = truth_img_valid.contiguous()
    valid = valid.view(torch.uint8) if valid.dtype == torch.bool else (valid != 0).view(torch.uint8)
    return _MaskLoss.apply(coeff, proto, pos_list, n_pos, match.match_index, _lib.f32c(truth_box), seg, valid)


def loss(prediction, truth, config):
    """``loss(prediction, truth, config) -> (total, (classification, box, mask))`` — loss.py:8-125, the reference's
    training entry point (yolact/scripts/train.py:246), same arguments and return value."""
    classification, box_encoding, mask_coeff, anchor, mask_prototype = prediction
    truth_valid, truth_classification, truth_box, truth_seg_map, truth_img_valid = truth
    match = match_anchors(anchor, truth_box, truth_valid, config)
    cb = class_box_loss(classification, box_encoding, match, truth_classification, config)
    ml = mask_loss(mask_coeff, mask_prototype, match, cb.pos_list, cb.n_pos, truth_box, truth_seg_map, truth_img_valid)
    total = cb.classification_loss + cb.box_loss + ml
    return total, (cb.classification_loss, cb.box_loss, ml)
