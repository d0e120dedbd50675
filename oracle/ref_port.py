"""CPU oracle — TEST INFRASTRUCTURE ONLY.

A restatement, in plain torch-CPU fp32 tensor ops, of the TAUV-Vision detection-head hot path
(scope rows a2-a19 of SURVEY.md section 8).  It exists to CHECK the CUDA path.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may
import it; nothing under ``tauv-vision_b200/`` does, and the product has no CPU path.

Why torch-CPU rather than numpy: the reference *is* a sequence of ATen calls, and its rounding
(sigmoid = 1/(1+exp(-x)) with SLEEF exp, softmax reduction order, int64->fp32 promotion rules) is
defined by ATen.  Re-using the same library for the arithmetic makes this port bit-identical to
the reference on CPU, so the only thing it adds is what the reference leaves undefined: a
canonical tie order — (score desc, flat index asc) for top-k, (confidence desc, prior index asc)
for NMS — which is the order the reference's own known-answer check asserts (decode.py:327-339).

Parity pinned: yes.  ``tests/golden/make_golden.py`` runs the REAL reference (imported from
/root/reference/src in the build container) on seeded inputs and commits its outputs under
``tests/golden/``; ``tests/test_oracle_vs_golden.py`` checks every function here against them,
plus the reference's two in-module KATs (decode.py:327-339, yolact boxes.py:106-117).

Each function cites the reference lines (under /root/reference/src/tauv_vision/) it follows.
"""
from __future__ import annotations

from dataclasses import dataclass
from math import floor, pi, sqrt
from typing import Optional, Tuple

import torch
import torch.nn.functional as F

# ------------------------------------------------------------------------------------------------
# CenterNet decode                                              centernet/model/decode.py
# ------------------------------------------------------------------------------------------------


def heatmap_nms(heatmap: torch.Tensor, kernel_size: int) -> torch.Tensor:
    """decode.py:239-252 — keep cells equal to their k x k neighbourhood max, zero the rest."""
    assert kernel_size >= 1 and kernel_size % 2 == 1
    pooled = F.max_pool2d(heatmap, kernel_size, stride=1, padding=(kernel_size - 1) // 2)
    return (pooled == heatmap).to(torch.float32) * heatmap


def stable_topk(scores: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Row-wise top-k with the canonical order (value desc, index asc).  scores [B, n]."""
    B, n = scores.shape
    if k > n:
        raise RuntimeError("selected index k out of range")
    vals = torch.empty((B, k), dtype=scores.dtype)
    idxs = torch.empty((B, k), dtype=torch.int64)
    for b in range(B):
        row = scores[b]
        kth = torch.topk(row, k).values[-1]
        cand = torch.nonzero(row >= kth, as_tuple=False).flatten()  # ascending index
        order = torch.sort(row[cand], descending=True, stable=True).indices[:k]
        idxs[b] = cand[order]
        vals[b] = row[idxs[b]]
    return vals, idxs


def heatmap_detect(heatmap: torch.Tensor, n_detections: int, canonical: bool = True):
    """decode.py:255-279 — joint top-k over C*H*W; (index [B,k,2] (y,x), label [B,k], score [B,k]).
    canonical=False uses a bare torch.topk exactly like the reference (arbitrary tie order) — that is the
    variant bench.py times, so the CPU baseline carries no cost the reference does not have."""
    B, C, H, W = heatmap.shape
    if canonical:
        score, flat = stable_topk(heatmap.reshape(B, -1), n_detections)
    else:
        score, flat = torch.topk(heatmap.reshape(B, -1), n_detections)
    label = torch.div(flat, H * W, rounding_mode="floor")
    rem = flat - label * (H * W)
    index = torch.stack((torch.div(rem, W, rounding_mode="floor"), rem % W), dim=-1)
    return index, label, score


def depth_decode(d: torch.Tensor) -> torch.Tensor:
    """decode.py:319-324."""
    return (1 / torch.sigmoid(d)) - 1


def angle_get_bins(bin_overlap: float):
    """decode.py:282-288."""
    return (pi / 2, -bin_overlap / 2, pi + bin_overlap / 2), (-pi / 2, -pi - bin_overlap / 2, bin_overlap / 2)


def angle_decode(predicted_bin, predicted_offset, theta_range: float, bin_overlap: float):
    """decode.py:291-316."""
    (c0, _, _), (c1, _, _) = angle_get_bins(bin_overlap)
    s0 = F.softmax(predicted_bin[:, :, 0:2], dim=-1)[:, :, 1]
    s1 = F.softmax(predicted_bin[:, :, 2:4], dim=-1)[:, :, 1]
    a0 = c0 + torch.atan2(predicted_offset[:, :, 0], predicted_offset[:, :, 1])
    a1 = c1 + torch.atan2(predicted_offset[:, :, 2], predicted_offset[:, :, 3])
    ang = torch.where(s1 > s0, a1, a0) % (2 * pi)
    return ang * (theta_range / (2 * pi))


@dataclass
class Packed:
    index: torch.Tensor   # [B,k,2] i64
    label: torch.Tensor   # [B,k] i64
    score: torch.Tensor   # [B,k] f32
    yx: torch.Tensor      # [B,k,2] f64
    hw: torch.Tensor      # [B,k,2] f32
    depth: Optional[torch.Tensor]  # [B,k] f32
    count: torch.Tensor   # [B] i32


def _gather_hw(t: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    """t [B,H,W,ch] (any strides), index [B,k,2] -> [B,k,ch]."""
    B = t.shape[0]
    bi = torch.arange(B).unsqueeze(1).expand(-1, index.shape[1])
    return t[bi, index[..., 0], index[..., 1]]


def decode_packed(heatmap_logits, size, offset, depth, downsample_ratio: int, in_h: int, in_w: int,
                  n_detections: int, score_threshold: float, canonical: bool = True) -> Packed:
    """decode.py:179-236 with the per-detection Python loop expressed as gathers.
    y = (ratio*iy + offset_y)/in_h in float64 (the reference computes these in Python floats, :214-215);
    count = entries before the first score < threshold (:208-209, fp32 compare)."""
    hm = heatmap_nms(torch.sigmoid(heatmap_logits), 3)
    index, label, score = heatmap_detect(hm, n_detections, canonical)
    off = _gather_hw(offset, index).to(torch.float64)
    yx = torch.stack(((downsample_ratio * index[..., 0].to(torch.float64) + off[..., 0]) / in_h,
                      (downsample_ratio * index[..., 1].to(torch.float64) + off[..., 1]) / in_w), dim=-1)
    hw = _gather_hw(size, index)
    dep = None
    if depth is not None:
        d = depth if depth.dim() == 4 else depth.unsqueeze(-1)
        dep = _gather_hw(depth_decode(d), index)[..., 0]
    below = score < score_threshold
    first = torch.where(below.any(dim=1), below.to(torch.int64).argmax(dim=1), torch.full((score.shape[0],), score.shape[1]))
    return Packed(index, label, score, yx, hw, dep, first.to(torch.int32))


def decode_keypoints_packed(heatmap_logits, size, depth, out_h: int, out_w: int, n_detections: int,
                            score_threshold: float) -> Packed:
    """Object half of decode_keypoints (decode.py:56-58, :65, :84-91): y = iy/out_h (fp32 divide),
    no offset, depth = 1/sigmoid(d)."""
    hm = heatmap_nms(torch.sigmoid(heatmap_logits), 3)
    index, label, score = heatmap_detect(hm, n_detections)
    yx = torch.stack(((index[..., 0] / out_h), (index[..., 1] / out_w)), dim=-1).to(torch.float64)
    hw = _gather_hw(size, index)
    dep = None
    if depth is not None:
        d = depth if depth.dim() == 4 else depth.unsqueeze(-1)
        dep = _gather_hw(1 / torch.sigmoid(d), index)[..., 0]
    below = score < score_threshold
    first = torch.where(below.any(dim=1), below.to(torch.int64).argmax(dim=1), torch.full((score.shape[0],), score.shape[1]))
    return Packed(index, label, score, yx, hw, dep, first.to(torch.int32))


# ------------------------------------------------------------------------------------------------
# CenterNet target encode                                        centernet/model/loss.py
# ------------------------------------------------------------------------------------------------


def gaussian_splat(h: int, w: int, cy: int, cx: int, sigma: float) -> torch.Tensor:
    """Missing from the reference snapshot; semantics from loss.py:64-67 and the call sites
    decode.py:328-332 / tests/centernet_square_detection.py:108-112."""
    y, x = torch.meshgrid(torch.arange(0, h), torch.arange(0, w), indexing="ij")
    return torch.exp(-((x - cx) ** 2 + (y - cy) ** 2) / (2 * sigma ** 2))


def generate_heatmap(valid, label, center, n_labels: int, out_h: int, out_w: int, in_h: int, in_w: int,
                     downsample_ratio: int, sigma: float) -> torch.Tensor:
    """loss.py:31-72 — per valid object, running max of a full-plane Gaussian at the floored centre."""
    B, n = valid.shape
    out = torch.zeros((B, n_labels, out_h, out_w), dtype=torch.float32)
    y, x = torch.meshgrid(torch.arange(0, out_h), torch.arange(0, out_w), indexing="ij")
    if sigma < 0.1:
        sigma = 0.1
    for b in range(B):
        for o in range(n):
            if not valid[b, o]:
                continue
            cy = floor(center[b, o, 0] * in_h / downsample_ratio)
            cx = floor(center[b, o, 1] * in_w / downsample_ratio)
            g = torch.exp(-((x - cx) ** 2 + (y - cy) ** 2) / (2 * sigma ** 2))
            out[b, label[b, o]] = torch.maximum(out[b, label[b, o]], g)
    return torch.nan_to_num(out)


def focal_loss(prediction: torch.Tensor, truth: torch.Tensor, alpha: float, beta: float) -> torch.Tensor:
    """loss.py:302-317 — penalty-reduced pixelwise focal loss of CenterNet, elementwise; the caller sums it
    (loss.py:233-234).  Positives are the cells whose target is (close to) 1; N counts them over the whole batch."""
    p = torch.isclose(truth, torch.ones(1))
    N = torch.sum(p)
    loss_p = ((1 - prediction) ** alpha) * torch.log(torch.clamp(prediction, min=1e-4)) * p.float()
    loss_n = ((1 - truth) ** beta) * (prediction ** alpha) * torch.log(torch.clamp(1 - prediction, min=1e-4)) * (1 - p.float())
    if N == 0:
        return -loss_p
    return -(loss_p + loss_n) / N


def heatmap_focal_loss(logits, valid, label, center, in_h: int, in_w: int, downsample_ratio: int, sigma: float,
                       alpha: float, beta: float):
    """loss.py:182 + :233-236 — the heatmap term of the CenterNet loss: render the Gaussian target, focal loss of the
    sigmoid of the logits against it, summed.  Returns (loss scalar, n_pos)."""
    B, C, H, W = logits.shape
    target = generate_heatmap(valid, label, center, C, H, W, in_h, in_w, downsample_ratio, sigma)
    loss = focal_loss(torch.sigmoid(logits), target, alpha, beta)
    return loss.sum(), int(torch.isclose(target, torch.ones(1)).sum())


def generate_keypoint_heatmap(kp_valid, kp_label, kp_center, kp_object_index, center, n_keypoints: int,
                              out_h: int, out_w: int, in_h: int, in_w: int, downsample_ratio: int,
                              sigma_heatmap: float, sigma_affinity: float):
    """loss.py:75-135 — keypoint heatmap, affinity weight, and the unit vector field from the owning
    object's centre (nearest object wins, strict '<' so the first instance wins ties)."""
    B, m = kp_valid.shape
    hm = torch.zeros((B, n_keypoints, out_h, out_w), dtype=torch.float32)
    wt = torch.zeros((B, n_keypoints, out_h, out_w), dtype=torch.float32)
    aff = torch.zeros((B, n_keypoints, 2, out_h, out_w), dtype=torch.float32)
    dist = torch.full((B, n_keypoints, out_h, out_w), float("inf"), dtype=torch.float32)
    y, x = torch.meshgrid(torch.arange(0, out_h), torch.arange(0, out_w), indexing="ij")
    grid = torch.stack((y / out_h, x / out_w), dim=0)
    for b in range(B):
        for i in range(m):
            if not kp_valid[b, i]:
                continue
            k = kp_label[b, i]
            cy = floor(kp_center[b, i, 0] * in_h / downsample_ratio)
            cx = floor(kp_center[b, i, 1] * in_w / downsample_ratio)
            neg_d2 = -((x - cx) ** 2 + (y - cy) ** 2)
            hm[b, k] = torch.maximum(hm[b, k], torch.exp(neg_d2 / (2 * sigma_heatmap ** 2)))
            wt[b, k] = torch.maximum(wt[b, k], torch.exp(neg_d2 / (2 * sigma_affinity ** 2)))
            disp = grid - center[b, kp_object_index[b, i]].unsqueeze(1).unsqueeze(2)
            disp = torch.nan_to_num(disp, 0)
            d = torch.nan_to_num(torch.sqrt(disp[0] ** 2 + disp[1] ** 2), 1)
            unit = disp / d
            aff[b, k] = torch.where(d < dist[b, k], unit, aff[b, k])
            dist[b, k] = torch.min(dist[b, k], d)
    return torch.nan_to_num(hm), torch.nan_to_num(wt), torch.nan_to_num(aff)


def keypoint_affinity_loss(pred_affinity, kp_valid, kp_label, kp_center, kp_object_index, center, out_h: int, out_w: int,
                           in_h: int, in_w: int, downsample_ratio: int, sigma_heatmap: float, sigma_affinity: float):
    """loss.py:244-246 before the lambda — weighted squared error between the predicted keypoint-affinity field
    [B,Kp,2,H,W] and generate_keypoint_heatmap's unit-vector field, weighted by its affinity weight."""
    _, weight, target = generate_keypoint_heatmap(kp_valid, kp_label, kp_center, kp_object_index, center,
                                                  pred_affinity.shape[1], out_h, out_w, in_h, in_w, downsample_ratio,
                                                  sigma_heatmap, sigma_affinity)
    return (weight.unsqueeze(2) * F.mse_loss(pred_affinity, target, reduction="none")).sum()


def out_index_for_position(position, in_h: int, in_w: int, downsample_ratio: int, out_h: int, out_w: int):
    """loss.py:138-142."""
    return torch.stack((
        torch.clamp(((position[..., 0] * in_h) / downsample_ratio).to(torch.long), 0, out_h - 1),
        torch.clamp(((position[..., 1] * in_w) / downsample_ratio).to(torch.long), 0, out_w - 1),
    ), dim=-1)


def offset_target(center, in_h: int, in_w: int, downsample_ratio: int):
    """loss.py:263-264."""
    pix = center * torch.tensor((in_h, in_w), dtype=torch.float32)
    return pix - downsample_ratio * (pix / downsample_ratio).to(torch.long)


# ------------------------------------------------------------------------------------------------
# YOLACT                                                         yolact/model/{anchors,boxes,nms,masks,loss}.py
# ------------------------------------------------------------------------------------------------


def get_anchor(fpn_i: int, fpn_size, anchor_scales, anchor_aspect_ratios, in_h: int, in_w: int) -> torch.Tensor:
    """anchors.py:9-41 — [1, A*H*W, 4] (y,x,h,w), aspect-major inside the level."""
    H, W = fpn_size
    cy = (torch.arange(0, H) + 0.5) / H
    cx = (torch.arange(0, W) + 0.5) / W
    gy, gx = torch.meshgrid(cy, cx, indexing="ij")
    gy, gx = gy.flatten(), gx.flatten()
    rows = []
    in_size = (in_h + in_w) / 2
    for ar in anchor_aspect_ratios:
        h = (anchor_scales[fpn_i] / in_size) * sqrt(ar)
        w = (anchor_scales[fpn_i] / in_size) / sqrt(ar)
        rows.append(torch.stack((gy, gx, torch.full_like(gy, h), torch.full_like(gx, w)), dim=-1))
    return torch.cat(rows, dim=0).unsqueeze(0)


def all_anchors(fpn_sizes, anchor_scales, anchor_aspect_ratios, in_h: int, in_w: int) -> torch.Tensor:
    """model.py:47-58 — the levels concatenated."""
    return torch.cat([get_anchor(i, s, anchor_scales, anchor_aspect_ratios, in_h, in_w)
                      for i, s in enumerate(fpn_sizes)], dim=1)


def box_to_corners(box):
    """boxes.py:15-27."""
    return torch.stack((box[..., 0] - box[..., 2] / 2, box[..., 1] - box[..., 3] / 2,
                        box[..., 0] + box[..., 2] / 2, box[..., 1] + box[..., 3] / 2), dim=-1)


def corners_to_box(c):
    """boxes.py:30-42."""
    return torch.stack(((c[..., 0] + c[..., 2]) / 2, (c[..., 1] + c[..., 3]) / 2,
                        c[..., 2] - c[..., 0], c[..., 3] - c[..., 1]), dim=-1)


def box_decode(enc, anchor, variances):
    """boxes.py:55-61."""
    return torch.cat((anchor[:, :, :2] + enc[:, :, :2] * variances[0] * anchor[:, :, 2:],
                      anchor[:, :, 2:] * torch.exp(enc[:, :, 2:] * variances[1])), -1)


def box_encode(box, anchor, variances):
    """boxes.py:45-52."""
    g_yx = (box[:, :, :2] - anchor[:, :, :2]) / (variances[0] * anchor[:, :, 2:])
    g_hw = torch.log(box[:, :, 2:] / anchor[:, :, 2:]) / variances[1]
    return torch.cat((g_yx, g_hw), -1)


def iou_matrix(box_a, box_b):
    """boxes.py:64-85 — [Ba,Na,4] x [Bb,Nb,4] -> [B,Na,Nb]; areas from (h*w), union = (a+b)-inter."""
    ca, cb = box_to_corners(box_a), box_to_corners(box_b)
    y0 = torch.max(ca[:, :, 0].unsqueeze(2), cb[:, :, 0].unsqueeze(1))
    x0 = torch.max(ca[:, :, 1].unsqueeze(2), cb[:, :, 1].unsqueeze(1))
    y1 = torch.min(ca[:, :, 2].unsqueeze(2), cb[:, :, 2].unsqueeze(1))
    x1 = torch.min(ca[:, :, 3].unsqueeze(2), cb[:, :, 3].unsqueeze(1))
    inter = torch.clamp(y1 - y0, min=0) * torch.clamp(x1 - x0, min=0)
    area_a = box_a[:, :, 2] * box_a[:, :, 3]
    area_b = box_b[:, :, 2] * box_b[:, :, 3]
    return inter / ((area_a.unsqueeze(2) + area_b.unsqueeze(1)) - inter)


def nms_scores(classification):
    """nms.py:9-10 — max foreground softmax confidence per prior."""
    return torch.max(F.softmax(classification, dim=-1)[:, :, 1:], dim=-1).values


def nms_frame(score_row, box_row, top_k: int, iou_threshold: float, confidence_threshold: float):
    """nms.py:12-27 for one frame given its confidences: canonical (confidence desc, prior asc) order,
    upper-triangular IoU, column max, keep = (iou_max <= thr) & (conf >= thr)."""
    conf, idx = torch.sort(score_row, descending=True, stable=True)
    idx, conf = idx[:top_k], conf[:top_k]
    b = box_row[idx].unsqueeze(0)
    iou = torch.triu(iou_matrix(b, b), diagonal=1)
    iou_max = torch.max(iou, dim=1).values[0]
    keep = (iou_max <= iou_threshold) & (conf >= confidence_threshold)
    return idx[keep], idx, conf


def nms(classification, box, top_k: int, iou_threshold: float, confidence_threshold: float):
    """nms.py:7-29 — frame 0 only, like the reference."""
    return nms_frame(nms_scores(classification)[0], box[0], top_k, iou_threshold, confidence_threshold)[0]


def box_to_mask(box, img_size):
    """boxes.py:88-103 — inclusive crop on integer pixel coordinates."""
    yg = torch.arange(0, img_size[0], dtype=torch.float, device=box.device)
    xg = torch.arange(0, img_size[1], dtype=torch.float, device=box.device)
    yc, xc = torch.meshgrid(yg, xg, indexing="ij")
    b = box * torch.tensor([img_size[0], img_size[1], img_size[0], img_size[1]], device=box.device)
    left, right = b[1] - b[3] / 2, b[1] + b[3] / 2
    top, bottom = b[0] - b[2] / 2, b[0] + b[2] / 2
    return ((xc >= left) & (xc <= right) & (yc >= top) & (yc <= bottom)).float()


def mask_logits(mask_prototype, mask_coeff):
    """masks.py:13 — sum_p coeff[i,p]*proto[p], fp32, summed over p in index order like torch.sum(dim=0)
    on a [P,H,W] temporary."""
    out = torch.empty((mask_coeff.shape[0],) + tuple(mask_prototype.shape[1:]), dtype=torch.float32)
    for i in range(mask_coeff.shape[0]):
        out[i] = torch.sum(mask_coeff[i].unsqueeze(1).unsqueeze(2) * mask_prototype, dim=0)
    return out


def assemble_mask(mask_prototype, mask_coeff, box):
    """masks.py:8-21."""
    m = torch.sigmoid(mask_logits(mask_prototype, mask_coeff))
    if box is not None:
        for i in range(m.shape[0]):
            m[i] *= box_to_mask(box[i], m[i].shape)
    return m


def upsample_nearest_index(out_size: int, in_size: int) -> torch.Tensor:
    """Source index of every output index for ``F.interpolate(x, size)`` in its default 'nearest' mode (the call at
    yolact_node.py:131): min(floor(i * fp32(in/out)), in - 1), the product in fp32 (ATen UpSample.h,
    nearest_neighbor_compute_source_index; checked against F.interpolate itself in tests/test_oracle_vs_golden.py)."""
    scale = torch.tensor(in_size, dtype=torch.float32) / torch.tensor(out_size, dtype=torch.float32)
    i = torch.arange(out_size, dtype=torch.float32)
    return torch.clamp(torch.floor(i * scale).to(torch.int64), max=in_size - 1)


def masked_depth_mean(mask_prototype, mask_coeff, box, depth_mm):
    """yolact_node.py:102-103 (depth image: 0 -> NaN, millimetres -> metres in float64), :130-131 (assemble_mask, then
    nearest-neighbour resize to the camera resolution), :178 (nanmean of the depth where the mask is > 0.5).
    depth_mm: [Hi,Wi] integer tensor (mono16).  Returns (mean [n] float64 with NaN where nothing was averaged,
    count [n] int64 = readings averaged)."""
    hi, wi = depth_mm.shape
    d = depth_mm.to(torch.float64)
    d = torch.where(depth_mm == 0, torch.full_like(d, float("nan")), d) / 1000
    mask = assemble_mask(mask_prototype, mask_coeff, box)
    iy = upsample_nearest_index(hi, mask.shape[1])
    ix = upsample_nearest_index(wi, mask.shape[2])
    up = mask[:, iy][:, :, ix]
    sel = (up > 0.5) & ~torch.isnan(d).unsqueeze(0)
    count = sel.sum(dim=(1, 2))
    total = torch.where(sel, d.unsqueeze(0).expand_as(up), torch.zeros((), dtype=torch.float64)).sum(dim=(1, 2))
    mean = torch.where(count > 0, total / count.clamp(min=1), torch.full_like(total, float("nan")))
    return mean, count


def upsample_bilinear_taps(out_size: int, in_size: int):
    """Taps of ``F.interpolate(x, size, mode="bilinear")`` (align_corners=False; the call at evaluate_batch.py:101):
    source coordinate max(fp32(in/out) * (i + 0.5) - 0.5, 0), left tap its integer part, right tap one further unless
    that leaves the image, weights (1 - frac, frac), all in fp32 (ATen UpSample.h, area_pixel_compute_source_index).
    The multiply-subtract is ONE fused operation in the builds measured (x86 CPU and nvcc both contract it), which
    shows in the last bits of the weights: the weights torch uses, recovered by interpolating an identity matrix, are
    reproduced exactly by the fused form and not by the unfused one (tests/test_oracle_vs_golden.py).  The product of
    two fp32 numbers is exact in float64, so float64 arithmetic rounded once is that fused operation.
    Returns (i0, i1, w0, w1)."""
    scale = (torch.tensor(in_size, dtype=torch.float32) / torch.tensor(out_size, dtype=torch.float32)).to(torch.float64)
    i = torch.arange(out_size, dtype=torch.float64)
    s = torch.clamp((scale * (i + 0.5) - 0.5).to(torch.float32), min=0.0)
    i0 = torch.clamp(s.to(torch.int64), max=in_size - 1)
    i1 = i0 + (i0 < in_size - 1).to(torch.int64)
    w1 = s - i0.to(torch.float32)
    return i0, i1, 1.0 - w1, w1


def mask_upsampled(mask_prototype, mask_coeff, box, size, mode: str):
    """yolact_node.py:130-135 (mode "nearest") / evaluate_batch.py:100-101 (mode "bilinear"): assemble_mask, then
    F.interpolate to `size`.  Returns the resized fp32 masks [n, size[0], size[1]]."""
    mask = assemble_mask(mask_prototype, mask_coeff, box)
    if mode == "nearest":
        iy = upsample_nearest_index(size[0], mask.shape[1])
        ix = upsample_nearest_index(size[1], mask.shape[2])
        return mask[:, iy][:, :, ix]
    assert mode == "bilinear", mode
    y0, y1, wy0, wy1 = upsample_bilinear_taps(size[0], mask.shape[1])
    x0, x1, wx0, wx1 = upsample_bilinear_taps(size[1], mask.shape[2])
    top, bot = mask[:, y0], mask[:, y1]
    return (wy0.view(1, -1, 1) * (wx0 * top[:, :, x0] + wx1 * top[:, :, x1])
            + wy1.view(1, -1, 1) * (wx0 * bot[:, :, x0] + wx1 * bot[:, :, x1]))


def mask_binary(mask_prototype, mask_coeff, box, size, mode: str):
    """The resized mask thresholded as its consumers do (`mask_np > 0.5`, yolact_node.py:178; `mask = mask > 0.5`,
    evaluate_batch.py:102), as uint8."""
    return (mask_upsampled(mask_prototype, mask_coeff, box, size, mode) > 0.5).to(torch.uint8)


def match_anchors(anchor, truth_box, truth_valid, pos_thr: float, neg_thr: float, variances):
    """yolact/model/loss.py:16-22 (+ :62-66 box_encode of the matched truth, here dense over all
    priors).  Returns (match_index [B,N], match_iou [B,N], positive, negative, target [B,N,4])."""
    iou = iou_matrix(anchor, truth_box)
    match_iou, match_index = torch.max(iou * truth_valid.unsqueeze(1).float(), dim=2)
    positive = match_iou >= pos_thr
    negative = match_iou <= neg_thr
    B = truth_box.shape[0]
    matched = truth_box[torch.arange(B, device=truth_box.device).unsqueeze(1), match_index]  # [B,N,4]
    target = box_encode(matched, anchor.expand(B, -1, -1), variances)
    return match_index, match_iou, positive, negative, target


def yolact_class_box_loss(classification, box_encoding, anchor, truth_valid, truth_classification, truth_box,
                          pos_thr: float, neg_thr: float, variances, ratio: int):
    """yolact/model/loss.py:16-73 — anchor matching, then the classification term with hard-negative mining and the
    box term.  Returns (classification_loss, box_loss, selected [B,N] bool).  The reference's ``torch.topk`` over
    ``where(negative, -background_confidence, -inf)`` (:40-43) leaves the order of equal entries open; here the lower
    prior index wins (a stable ascending sort), which is also what the CUDA path does."""
    match_index, match_iou, positive, negative, _ = match_anchors(anchor, truth_box, truth_valid, pos_thr, neg_thr, variances)
    B, N = positive.shape
    cls_sums, box_sums, selected = [], [], torch.zeros((B, N), dtype=torch.bool, device=classification.device)
    for b in range(B):
        target_class = truth_classification[b, match_index[b]].clone()          # :27
        target_class[~positive[b]] = 0                                          # :28
        ce = F.cross_entropy(classification[b], target_class, reduction="none")  # :30-34
        k = int(ratio * positive[b].sum())                                      # :35-36
        background = F.softmax(classification[b], dim=-1)[:, 0]                 # :38
        key = torch.where(negative[b], background, torch.full_like(background, float("inf")))  # :40-43, ascending
        mined = torch.sort(key, stable=True).indices[:k].detach()
        sel = positive[b].clone()                                               # :48-50
        sel[mined] = True
        selected[b] = sel
        cls_sums.append((sel.float() * ce).sum())                               # :52
        pos = positive[b]
        tgt = box_encode(truth_box[b, match_index[b, pos]].unsqueeze(0), anchor[0, pos].unsqueeze(0), variances).squeeze(0)
        box_sums.append(F.smooth_l1_loss(box_encoding[b, pos], tgt, reduction="none").sum())   # :61-66
    P = positive.sum()
    cls_loss, box_loss = torch.stack(cls_sums).sum(), torch.stack(box_sums).sum()
    if P > 0:                                                                   # :54-57, :70-73
        cls_loss, box_loss = cls_loss / ((1 + ratio) * P), box_loss / P
    return cls_loss, box_loss, selected


def yolact_mask_loss(mask_coeff, mask_prototype, anchor, truth_valid, truth_box, truth_seg_map, truth_img_valid,
                     pos_thr: float, neg_thr: float, variances):
    """yolact/model/loss.py:75-121 — the mask term: for every positive prior, BCE of the assembled mask against the
    bilinearly resized mask of its matched truth, inside the truth box and the valid image region, over the resized
    truth mask's area; positives whose resized truth mask is empty are skipped."""
    match_index, _, positive, _, _ = match_anchors(anchor, truth_box, truth_valid, pos_thr, neg_thr, variances)
    B = positive.shape[0]
    size = mask_prototype.shape[-2:]
    total = torch.zeros((), device=mask_coeff.device)
    for b in range(B):
        for n in positive[b].nonzero().flatten().tolist():
            j = int(match_index[b, n])
            m = (mask_coeff[b, n].view(-1, 1, 1) * mask_prototype[b]).sum(dim=0)             # :82
            m = torch.clamp(torch.sigmoid(m), min=1e-4)                                     # :83-84
            t = F.interpolate((truth_seg_map[b] == j).float()[None, None], size, mode="bilinear")[0, 0]  # :86-91
            if t.sum() == 0:                                                                # :93-94
                continue
            bce = F.binary_cross_entropy(torch.clamp(m.reshape(-1), 1e-4, 1 - 1e-4),
                                         torch.clamp(t.reshape(-1), 1e-4, 1 - 1e-4), reduction="none")  # :96-100
            valid = F.interpolate(truth_img_valid[b].float()[None, None], size, mode="nearest")[0, 0]   # :102-106
            w = box_to_mask(truth_box[b, j], size) * valid                                  # :108-111
            total = total + (w.reshape(-1) * bce).sum() / t.sum()                           # :113
    P = positive.sum()
    return total / P if P > 0 else total                                                    # :117-120


def pack_head(levels, channels_per_prior: int, tanh: bool = False):
    """yolact/model/prediction_head.py:111-113 / :122-124 / :137-140 per level, then model.py:55-58's torch.cat along the
    prior axis: NCHW head outputs [B, A*C, H_l, W_l] -> [B, sum_l H_l*W_l*A, C]."""
    out = torch.cat([t.permute(0, 2, 3, 1).reshape(t.size(0), -1, channels_per_prior) for t in levels], dim=1)
    return torch.tanh(out) if tanh else out
