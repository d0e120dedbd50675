"""CenterNet training-target encoding — drop-in for the target-render half of
``tauv_vision.centernet.model.loss`` (/root/reference/src/tauv_vision/centernet/model/loss.py):
``generate_heatmap`` (:31-72), ``generate_keypoint_heatmap`` (:75-135),
``out_index_for_position`` (:138-142), the sub-pixel offset target (:263-264) and
``gaussian_splat`` (missing from the snapshot; call sites decode.py:328-332).

Of the loss arithmetic, the heatmap term — the only one that touches [B,C,H,W] tensors — is here too, fused with
its target render (``heatmap_focal_loss``, loss.py:182 + :233-236 + :302-317, with autograd); the per-object L1 /
angle / depth terms stay with the caller.  Kernels: csrc/gaussian_encode.cu.
"""
from __future__ import annotations

from typing import Tuple

import torch

from ... import _lib


def _u8(t: torch.Tensor) -> torch.Tensor:
    t = t.contiguous()
    return t.view(torch.uint8) if t.dtype == torch.bool else t.to(torch.uint8)


def _i64(t: torch.Tensor) -> torch.Tensor:
    return t.contiguous() if t.dtype == torch.int64 else t.to(torch.int64).contiguous()


def generate_heatmap(truth, model_config, train_config, object_config, out: torch.Tensor = None) -> torch.Tensor:
    """[B, n_labels, out_h, out_w] Gaussian class heatmap   — reference loss.py:31-72.
    ``out`` (optional, beyond the reference signature): a contiguous fp32 tensor of that shape to overwrite."""
    dev = _lib.require_cuda(truth.valid, truth.label, truth.center)
    B, n = truth.valid.shape
    C = int(object_config.n_labels)
    H, W = int(model_config.out_h), int(model_config.out_w)
    if out is None:
        out = torch.empty((B, C, H, W), dtype=torch.float32, device=dev)
    elif tuple(out.shape) != (B, C, H, W) or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError("`out` must be a contiguous fp32 [B, n_labels, out_h, out_w] tensor")
    valid, label, center = _u8(truth.valid), _i64(truth.label), _lib.f32c(truth.center)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().tauv_gaussian_encode(
            _lib.u8ptr(valid), _lib.i64ptr(label), _lib.fptr(center), B, n, C, H, W,
            int(model_config.in_h), int(model_config.in_w), int(model_config.downsample_ratio),
            float(train_config.keypoint_heatmap_sigma), _lib.fptr(out), _lib.stream_ptr(dev)))
    return out


def generate_keypoint_heatmap(truth, model_config, train_config, object_config
                              ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(heatmap [B,Kp,H,W], affinity_weight [B,Kp,H,W], affinity [B,Kp,2,H,W])   — loss.py:75-135."""
    dev = _lib.require_cuda(truth.keypoint_valid, truth.keypoint_label, truth.keypoint_center,
                            truth.keypoint_object_index, truth.center)
    B, m = truth.keypoint_valid.shape
    n_obj = truth.center.shape[1]
    Kp = int(object_config.n_keypoints)
    H, W = int(model_config.out_h), int(model_config.out_w)
    heatmap = torch.empty((B, Kp, H, W), dtype=torch.float32, device=dev)
    weight = torch.empty((B, Kp, H, W), dtype=torch.float32, device=dev)
    affinity = torch.empty((B, Kp, 2, H, W), dtype=torch.float32, device=dev)
    kv, kl = _u8(truth.keypoint_valid), _i64(truth.keypoint_label)
    kc, ko = _lib.f32c(truth.keypoint_center), _i64(truth.keypoint_object_index)
    center = _lib.f32c(truth.center)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().tauv_keypoint_encode(
            _lib.u8ptr(kv), _lib.i64ptr(kl), _lib.fptr(kc), _lib.i64ptr(ko), _lib.fptr(center),
            B, m, n_obj, Kp, H, W, int(model_config.in_h), int(model_config.in_w),
            int(model_config.downsample_ratio), float(train_config.keypoint_heatmap_sigma),
            float(train_config.keypoint_affinity_sigma), _lib.fptr(heatmap), _lib.fptr(weight),
            _lib.fptr(affinity), _lib.stream_ptr(dev)))
    return heatmap, weight, affinity


def _index_offset(position: torch.Tensor, model_config, want_offset: bool):
    dev = _lib.require_cuda(position)
    pos = _lib.f32c(position)
    n = pos.numel() // 2
    index = torch.empty(pos.shape, dtype=torch.int64, device=dev)
    offset = torch.empty(pos.shape, dtype=torch.float32, device=dev) if want_offset else None
    with torch.cuda.device(dev):
        _lib.check(_lib.load().tauv_out_index_offset(
            _lib.fptr(pos), n, int(model_config.in_h), int(model_config.in_w), int(model_config.downsample_ratio),
            int(model_config.out_h), int(model_config.out_w), _lib.i64ptr(index), _lib.fptr(offset),
            _lib.stream_ptr(dev)))
    return index, offset


def out_index_for_position(position: torch.Tensor, model_config) -> torch.Tensor:
    """[B,n,2] normalised (y,x) -> clamped output-grid cell [B,n,2] i64   — loss.py:138-142."""
    return _index_offset(position, model_config, False)[0]


def offset_target(center: torch.Tensor, model_config) -> torch.Tensor:
    """Sub-pixel regression target ``pix - ratio*trunc(pix/ratio)``   — loss.py:263-264."""
    return _index_offset(center, model_config, True)[1]


def gaussian_splat(h: int, w: int, cy: int, cx: int, sigma: float, device=None) -> torch.Tensor:
    """[h,w] plane exp(-((x-cx)^2+(y-cy)^2)/(2 sigma^2)); the helper the reference's own KAT
    (decode.py:327-339) and tests/centernet_square_detection.py:108-112 import."""
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if dev.type != "cuda":
        raise RuntimeError("tauv_vision_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
    out = torch.empty((int(h), int(w)), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().tauv_gaussian_splat(int(h), int(w), int(cy), int(cx), float(sigma), _lib.fptr(out),
                                                   _lib.stream_ptr(dev)))
    return out


def focal_loss(prediction: torch.Tensor, truth: torch.Tensor, alpha: float, beta: float) -> torch.Tensor:
    """Elementwise penalty-reduced focal loss — reference loss.py:302-317, the same tensor expressions (any device;
    the reference's callers sum it).  ``heatmap_focal_loss`` is the fused form of
    ``focal_loss(sigmoid(logits), generate_heatmap(...)).sum()``."""
    p = torch.isclose(truth, torch.ones(1, device=truth.device))
    N = torch.sum(p)
    loss_p = ((1 - prediction) ** alpha) * torch.log(torch.clamp(prediction, min=1e-4)) * p.float()
    loss_n = ((1 - truth) ** beta) * (prediction ** alpha) * torch.log(torch.clamp(1 - prediction, min=1e-4)) * (1 - p.float())
    if N == 0:
        return -loss_p
    return -(loss_p + loss_n) / N


class _HeatmapFocalLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, valid, label, center, geom, sigma, alpha, beta):
        dev = logits.device
        B, C, H, W = logits.shape
        n = valid.shape[1]
        lib = _lib.load()
        sums = torch.empty((B, 2), dtype=torch.float64, device=dev)
        pos = torch.empty((B,), dtype=torch.int64, device=dev)
        ws = _lib.workspace(dev, lib.tauv_centernet_focal_loss_workspace_bytes(B, C, H, W))
        with torch.cuda.device(dev):
            _lib.check(lib.tauv_centernet_focal_loss(
                _lib.fptr(logits), _lib.u8ptr(valid), _lib.i64ptr(label), _lib.fptr(center), B, n, C, H, W, *geom,
                float(sigma), float(alpha), float(beta), _lib.dptr(sums), _lib.i64ptr(pos), ws.data_ptr(), ws.numel(),
                _lib.stream_ptr(dev)))
        loss = torch.empty((), dtype=torch.float32, device=dev)
        n_pos = torch.empty((1,), dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):   # the normalisation (loss.py:313-317), one small launch
            _lib.check(lib.tauv_centernet_focal_loss_reduce(_lib.dptr(sums), _lib.i64ptr(pos), B, _lib.fptr(loss),
                                                            _lib.i64ptr(n_pos), _lib.stream_ptr(dev)))
        ctx.save_for_backward(logits, valid, label, center, n_pos)
        ctx.meta = (geom, float(sigma), float(alpha), float(beta))
        ctx.mark_non_differentiable(n_pos)
        return loss, n_pos

    @staticmethod
    def backward(ctx, grad_loss, _grad_n):
        logits, valid, label, center, n_pos = ctx.saved_tensors
        geom, sigma, alpha, beta = ctx.meta
        dev = logits.device
        B, C, H, W = logits.shape
        grad = torch.empty_like(logits)
        go = grad_loss.to(torch.float32).contiguous().reshape(1)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().tauv_centernet_focal_loss_backward(
                _lib.fptr(logits), _lib.u8ptr(valid), _lib.i64ptr(label), _lib.fptr(center), B, valid.shape[1], C, H, W,
                *geom, sigma, alpha, beta, _lib.i64ptr(n_pos), _lib.fptr(go), _lib.fptr(grad), _lib.stream_ptr(dev)))
        return grad, None, None, None, None, None, None, None


def heatmap_focal_loss(heatmap_logits: torch.Tensor, truth, model_config, train_config, return_n_pos: bool = False):
    """``focal_loss(F.sigmoid(prediction.heatmap), generate_heatmap(truth, ...), a, b).sum()`` — the heatmap term of the
    reference's loss (loss.py:182, :233-236) — in one pass over the logits, without materialising the target, with
    autograd (one more pass writes the gradient).  ``heatmap_logits`` is ``prediction.heatmap`` (pre-sigmoid, [B,C,H,W]);
    the class count comes from its shape.  Returns the fp32 scalar (and the number of positive cells N if asked)."""
    dev = _lib.require_cuda(heatmap_logits, truth.valid, truth.label, truth.center)
    B, C, H, W = heatmap_logits.shape
    if (H, W) != (int(model_config.out_h), int(model_config.out_w)):
        raise ValueError(f"heatmap is {H}x{W}, the model's output grid is {model_config.out_h}x{model_config.out_w}")
    alpha, beta = float(train_config.heatmap_focal_loss_a), float(train_config.heatmap_focal_loss_b)
    sigma = float(train_config.keypoint_heatmap_sigma)
    logits = _lib.f32c(heatmap_logits)
    valid, label, center = _u8(truth.valid), _i64(truth.label), _lib.f32c(truth.center)
    fused = W % 4 == 0 and logits.data_ptr() % 16 == 0 and valid.shape[1] <= 32 and center.data_ptr() % 8 == 0
    if not fused:
        # shapes the fused kernel does not take: the same arithmetic from the rendered target, still on the GPU
        target = generate_heatmap(truth, model_config, train_config, type("O", (), {"n_labels": C}))
        loss = focal_loss(torch.sigmoid(logits), target, alpha, beta).sum()
        n_pos = torch.isclose(target, torch.ones(1, device=dev)).sum().reshape(1)
        return (loss, n_pos) if return_n_pos else loss
    geom = (int(model_config.in_h), int(model_config.in_w), int(model_config.downsample_ratio))
    loss, n_pos = _HeatmapFocalLoss.apply(logits, valid, label, center, geom, sigma, alpha, beta)
    return (loss, n_pos) if return_n_pos else loss


def keypoint_heatmap_focal_loss(keypoint_heatmap_logits: torch.Tensor, truth, model_config, train_config,
                                return_n_pos: bool = False):
    """``focal_loss(F.sigmoid(prediction.keypoint_heatmap), generate_keypoint_heatmap(truth, ...)[0], a, b).sum()`` — the
    keypoint-heatmap term of the reference's loss before its lambda (loss.py:238-240).  The keypoint heatmap target is
    rendered exactly like the object heatmap (loss.py:113-116 vs :64-67: same sigma, same floor of the centre, maximum
    over the instances of a label), so this is the fused pass of ``heatmap_focal_loss`` over the keypoint instances;
    the channel count comes from the logits' shape."""
    if float(train_config.keypoint_heatmap_sigma) < 0.1:
        raise ValueError("keypoint_heatmap_sigma < 0.1: generate_heatmap floors sigma at 0.1 (loss.py:60-62), "
                         "generate_keypoint_heatmap does not (loss.py:115)")
    kp = type("KeypointTruth", (), {"valid": truth.keypoint_valid, "label": truth.keypoint_label,
                                    "center": truth.keypoint_center})
    return heatmap_focal_loss(keypoint_heatmap_logits, kp, model_config, train_config, return_n_pos=return_n_pos)


class _KeypointAffinityLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, kv, kl, kc, ko, center, geom, sigma_a):
        dev = pred.device
        B, Kp, _, H, W = pred.shape
        lib = _lib.load()
        partial = torch.empty((lib.tauv_keypoint_affinity_loss_partials(B, Kp, H, W),), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.tauv_keypoint_affinity_loss(
                _lib.fptr(pred), _lib.u8ptr(kv), _lib.i64ptr(kl), _lib.fptr(kc), _lib.i64ptr(ko), _lib.fptr(center), B,
                kv.shape[1], center.shape[1], Kp, H, W, *geom, float(sigma_a), _lib.dptr(partial), _lib.stream_ptr(dev)))
        ctx.save_for_backward(pred, kv, kl, kc, ko, center)
        ctx.meta = (geom, float(sigma_a))
        return partial.sum().to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        pred, kv, kl, kc, ko, center = ctx.saved_tensors
        geom, sigma_a = ctx.meta
        dev = pred.device
        B, Kp, _, H, W = pred.shape
        grad = torch.empty_like(pred)
        go = g.to(torch.float32).contiguous().reshape(1)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().tauv_keypoint_affinity_loss_backward(
                _lib.fptr(pred), _lib.u8ptr(kv), _lib.i64ptr(kl), _lib.fptr(kc), _lib.i64ptr(ko), _lib.fptr(center), B,
                kv.shape[1], center.shape[1], Kp, H, W, *geom, sigma_a, _lib.fptr(go), _lib.fptr(grad),
                _lib.stream_ptr(dev)))
        return grad, None, None, None, None, None, None, None


def keypoint_affinity_loss(keypoint_affinity: torch.Tensor, truth, model_config, train_config) -> torch.Tensor:
    """``(keypoint_affinity_weight.unsqueeze(2) * F.mse_loss(prediction.keypoint_affinity, keypoint_affinity,
    reduction="none")).sum()`` — the keypoint-affinity term of the reference's loss before its lambda (loss.py:244-246)
    — with the weight and the target field of ``generate_keypoint_heatmap`` computed on the fly (1.3 GB of targets at
    Kp = 80 are never written; planes without instances are not even read), with autograd.
    ``keypoint_affinity`` is ``prediction.keypoint_affinity`` [B,Kp,2,H,W]."""
    dev = _lib.require_cuda(keypoint_affinity, truth.keypoint_valid, truth.keypoint_label, truth.keypoint_center,
                            truth.keypoint_object_index, truth.center)
    if keypoint_affinity.dim() != 5 or keypoint_affinity.shape[2] != 2:
        raise ValueError(f"keypoint_affinity must be [B,Kp,2,H,W]; got {tuple(keypoint_affinity.shape)}")
    B, Kp, _, H, W = keypoint_affinity.shape
    if (H, W) != (int(model_config.out_h), int(model_config.out_w)):
        raise ValueError(f"affinity field is {H}x{W}, the model's output grid is {model_config.out_h}x{model_config.out_w}")
    pred = _lib.f32c(keypoint_affinity)
    kv, kl = _u8(truth.keypoint_valid), _i64(truth.keypoint_label)
    kc, ko, center = _lib.f32c(truth.keypoint_center), _i64(truth.keypoint_object_index), _lib.f32c(truth.center)
    m = kv.shape[1]
    if not (W % 4 == 0 and pred.data_ptr() % 16 == 0 and 1 <= m <= 128 and center.shape[1] > 0):
        # shapes the fused kernel does not take: the same arithmetic from the rendered targets, still on the GPU
        _, weight, target = generate_keypoint_heatmap(truth, model_config, train_config, type("O", (), {"n_keypoints": Kp}))
        return (weight.unsqueeze(2) * (pred - target) ** 2).sum()
    del dev
    geom = (int(model_config.in_h), int(model_config.in_w), int(model_config.downsample_ratio))
    return _KeypointAffinityLoss.apply(pred, kv, kl, kc, ko, center, geom, float(train_config.keypoint_affinity_sigma))


class _GatherAtObjects(torch.autograd.Function):
    @staticmethod
    def forward(ctx, head, index):
        dev = head.device
        B, H, W, C = head.shape
        n = index.shape[1]
        out = torch.empty((B, n, C), dtype=torch.float32, device=dev)
        sb, sy, sx, sc = head.stride()
        with torch.cuda.device(dev):
            _lib.check(_lib.load().tauv_gather_at(_lib.fptr(head), sb, 0, sc, sy, sx, C, _lib.i64ptr(index), None, B, n,
                                                  _lib.fptr(out), _lib.stream_ptr(dev)))
        ctx.save_for_backward(index)
        ctx.shape = (B, H, W, C)
        return out

    @staticmethod
    def backward(ctx, grad):
        index, = ctx.saved_tensors
        B, H, W, C = ctx.shape
        dev = grad.device
        g = grad.to(torch.float32).contiguous()
        dst = torch.zeros((B, C, H, W), dtype=torch.float32, device=dev).permute(0, 2, 3, 1)  # the heads' own memory layout
        sb, sy, sx, sc = dst.stride()
        with torch.cuda.device(dev):
            _lib.check(_lib.load().tauv_scatter_add_at(_lib.fptr(g), _lib.i64ptr(index), B, index.shape[1], C, _lib.fptr(dst),
                                                       sb, sc, sy, sx, _lib.stream_ptr(dev)))
        return dst, None


def gather_at_objects(head: torch.Tensor, out_index: torch.Tensor) -> torch.Tensor:
    """``out[b, o] = head[b, out_index[b, o, 0], out_index[b, o, 1]]`` for every frame and object in one launch — the
    Python double loop of the reference's loss (loss.py:196-227: prediction_size, prediction_offset, the angle bins and
    offsets, prediction_depth), with autograd (objects that share a cell add up, in object order).  ``head`` is one of
    the prediction's ``[B,H,W,C]`` views (``size``, ``offset``, ``roll_bin`` ... ; read through its strides, never made
    contiguous) or ``[B,H,W]``; ``out_index`` is ``out_index_for_position(truth.center, model_config)`` [B,n,2].
    Returns ``[B,n,C]`` (``[B,n]`` for a ``[B,H,W]`` head)."""
    _lib.require_cuda(head, out_index)
    squeeze = head.dim() == 3
    h = head.unsqueeze(-1) if squeeze else head
    if h.dim() != 4 or out_index.dim() != 3 or out_index.shape[0] != h.shape[0] or out_index.shape[2] != 2:
        raise ValueError(f"head must be [B,H,W,C] and out_index [B,n,2]; got {tuple(head.shape)}, {tuple(out_index.shape)}")
    if h.dtype != torch.float32:
        h = h.float()
    out = _GatherAtObjects.apply(h, _i64(out_index))
    return out.squeeze(-1) if squeeze else out
