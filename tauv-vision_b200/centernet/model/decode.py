"""CenterNet head decode — drop-in for ``tauv_vision.centernet.model.decode``.

Same names, positional order and error behaviour as the reference
(/root/reference/src/tauv_vision/centernet/model/decode.py): ``heatmap_nms`` (:239-252),
``heatmap_detect`` (:255-279), ``decode`` (:179-236), ``decode_keypoints`` (:51-176),
``angle_get_bins`` (:282-288), ``angle_decode`` (:291-316), ``depth_decode`` (:319-324) and the
``Detection`` / ``KeypointDetection`` records (:16-48).  All tensor work happens in
libtauv_b200 (csrc/centernet_decode.cu); Python only builds the per-frame lists the reference
signature demands, after ONE device->host copy of the packed results (the reference does
~8 blocking scalar reads per detection).

Tie order (the reference leaves it to torch.topk): score descending, then flat index ascending —
the order the reference's own known-answer check (decode.py:327-339) asserts.
"""
from __future__ import annotations

from dataclasses import dataclass
from math import atan2, pi
from typing import Any, List, Optional, Tuple

import numpy as np
import torch

from ... import _lib

_NP_DTYPE = {torch.int64: np.int64, torch.float64: np.float64, torch.float32: np.float32, torch.int32: np.int32,
             torch.uint8: np.uint8}

TOPK_RAW = 0
TOPK_SIGMOID_PEAK = 1
BOX_DECODE = 0
BOX_KEYPOINTS = 1


@dataclass
class Detection:
    label: int
    score: float
    y: float
    x: float
    h: float
    w: float

    yaw: Optional[float] = None
    pitch: Optional[float] = None
    roll: Optional[float] = None

    depth: Optional[float] = None


@dataclass
class KeypointDetection:
    label: int
    score: float

    y: float
    x: float

    w: float
    h: float
    depth: float

    keypoints: List[Optional[Tuple[float, float, float]]]
    keypoint_scores: List[Optional[float]]
    keypoint_affinities: List[Optional[Tuple[float, float, float]]]

    cam_t_object: Any


@dataclass
class PackedDetections:
    """Device-resident result of one decode launch chain (nothing here has synchronised).

    index [B,k,2] i64 (y,x) · label [B,k] i64 · score [B,k] f32 · yx [B,k,2] f64 · hw [B,k,2] f32 ·
    depth [B,k] f32 or None · count [B] i32 = entries before the first score < threshold.
    """
    index: torch.Tensor
    label: torch.Tensor
    score: torch.Tensor
    yx: torch.Tensor
    hw: torch.Tensor
    depth: Optional[torch.Tensor]
    count: torch.Tensor
    extra: Optional[dict] = None             # further tensors carved from the same buffer (decode_keypoints)
    _storage: Optional[torch.Tensor] = None  # one uint8 buffer all the tensors above are views of (if allocated here)
    _layout: Optional[tuple] = None          # ((name, dtype, shape, byte offset, bytes), ...)

    _FIELDS = (("index", torch.int64, 8, 2), ("label", torch.int64, 8, 1), ("yx", torch.float64, 8, 2),
               ("score", torch.float32, 4, 1), ("hw", torch.float32, 4, 2), ("depth", torch.float32, 4, 1))

    @classmethod
    def allocate(cls, B: int, k: int, with_depth: bool, device, extra=()) -> "PackedDetections":
        """All outputs as typed views of ONE device buffer (8-byte fields first), so that ``to_host`` is one copy.
        ``extra``: further (name, dtype, shape) tensors to carve from the same buffer (``self.extra[name]``)."""
        layout, off = [], 0
        for name, dtype, isz, per in cls._FIELDS:
            if name == "depth" and not with_depth:
                continue
            layout.append((name, dtype, (B, k, per) if per > 1 else (B, k), off, B * k * per * isz))
            off += (B * k * per * isz + 15) // 16 * 16
        layout.append(("count", torch.int32, (B,), off, B * 4))
        off += (B * 4 + 15) // 16 * 16
        for name, dtype, shape in extra:
            nb = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
            layout.append((name, dtype, tuple(shape), off, nb))
            off += (nb + 15) // 16 * 16
        storage = torch.empty((off,), dtype=torch.uint8, device=device)
        views = {name: storage[o:o + nb].view(dtype).view(shape) for name, dtype, shape, o, nb in layout}
        return cls(index=views["index"], label=views["label"], score=views["score"], yx=views["yx"], hw=views["hw"],
                   depth=views.get("depth"), count=views["count"], extra={name: views[name] for name, _, _ in extra},
                   _storage=storage, _layout=tuple(layout))

    def to_host(self) -> dict:
        """One synchronising device->host transfer of everything."""
        if self._storage is not None:
            raw = self._storage.cpu().numpy()
            out = {"depth": None}
            for name, dtype, shape, o, nb in self._layout:
                out[name] = raw[o:o + nb].view(_NP_DTYPE[dtype]).reshape(shape)
            return out
        out = {k: getattr(self, k).cpu().numpy() for k in ("index", "label", "score", "yx", "hw", "count")}
        out["depth"] = self.depth.cpu().numpy() if self.depth is not None else None
        return out

    def to_lists(self) -> List[List[Detection]]:
        h = self.to_host()
        # (numpy -> Python scalars in bulk: .tolist() is several times faster than item-by-item conversion)
        label, score, yx, hw = h["label"].tolist(), h["score"].tolist(), h["yx"].tolist(), h["hw"].tolist()
        depth = h["depth"].tolist() if h["depth"] is not None else None
        frames = []
        for b, n in enumerate(h["count"].tolist()):
            lb, sb, yb, hb = label[b], score[b], yx[b], hw[b]
            db = depth[b] if depth is not None else None
            frames.append([Detection(label=lb[i], score=sb[i], y=yb[i][0], x=yb[i][1], h=hb[i][0], w=hb[i][1],
                                     depth=db[i] if db is not None else None) for i in range(n)])
        return frames


# ------------------------------------------------------------------------------------------------
# heatmap_nms / heatmap_detect
# ------------------------------------------------------------------------------------------------

def _as_heatmap(heatmap: torch.Tensor) -> torch.Tensor:
    if heatmap.dim() != 4:
        raise ValueError(f"heatmap must be [batch, n_heatmaps, h, w]; got {tuple(heatmap.shape)}")
    return _lib.f32c(heatmap)


def heatmap_nms(heatmap: torch.Tensor, kernel_size: int) -> torch.Tensor:
    """(max_pool2d(h, k, 1, (k-1)//2) == h).float() * h     — reference decode.py:239-252."""
    assert kernel_size >= 1 and kernel_size % 2 == 1
    dev = _lib.require_cuda(heatmap)
    hm = _as_heatmap(heatmap)
    out = torch.empty_like(hm)
    B, C, H, W = hm.shape
    with torch.cuda.device(dev):
        _lib.check(_lib.load().tauv_heatmap_nms(_lib.fptr(hm), _lib.fptr(out), B, C, H, W, int(kernel_size), 0,
                                                _lib.stream_ptr(dev)))
    return out


def _topk(hm: torch.Tensor, k: int, mode: int):
    dev = hm.device
    B, C, H, W = hm.shape
    lib = _lib.load()
    k = int(k)
    if k > C * H * W:
        raise RuntimeError(f"selected index k out of range (k={k} > {C * H * W})")
    index = torch.empty((B, k, 2), dtype=torch.int64, device=dev)
    label = torch.empty((B, k), dtype=torch.int64, device=dev)
    score = torch.empty((B, k), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        nbytes = lib.tauv_heatmap_topk_workspace_bytes(B, C, H, W, k)
        ws = _lib.workspace(dev, nbytes)
        _lib.check(lib.tauv_heatmap_topk(_lib.fptr(hm), B, C, H, W, k, mode, _lib.i64ptr(index), _lib.i64ptr(label),
                                         _lib.fptr(score), ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev)))
    return index, label, score


def heatmap_detect(heatmap: torch.Tensor, n_detections: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Joint top-k over C*H*W per frame -> (index [B,k,2], label [B,k], score [B,k])   — decode.py:255-279."""
    _lib.require_cuda(heatmap)
    return _topk(_as_heatmap(heatmap), n_detections, TOPK_RAW)


def heatmap_peaks(heatmap_logits: torch.Tensor, n_detections: int):
    """sigmoid -> heatmap_nms(3) -> heatmap_detect in one pass over the logits (decode.py:182-184)."""
    _lib.require_cuda(heatmap_logits)
    return _topk(_as_heatmap(heatmap_logits), n_detections, TOPK_SIGMOID_PEAK)


# ------------------------------------------------------------------------------------------------
# decode
# ------------------------------------------------------------------------------------------------

def _depth_view(depth: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if depth is None:
        return None
    if depth.dim() == 4:  # [B,H,W,1] permuted view (centernet.py:89)
        depth = depth[..., 0]
    return depth if depth.dtype == torch.float32 else depth.to(torch.float32)


def decode_packed(prediction, model_config, n_detections: int, score_threshold: float,
                  stage_events=None, out: Optional[PackedDetections] = None) -> PackedDetections:
    """Device part of ``decode``: two kernel launches, no synchronisation, permuted views taken as-is.

    ``stage_events`` (profiling hook): three ``torch.cuda.Event(enable_timing=True)`` recorded on the current
    stream before the tile kernel, between the two kernels, and after the merge kernel.
    ``out``: a PackedDetections from an earlier call with the same shapes, to be overwritten (no allocation)."""
    hm = prediction.heatmap
    dev = _lib.require_cuda(hm, prediction.size, prediction.offset, prediction.depth)
    hm = _as_heatmap(hm)
    size = prediction.size if prediction.size.dtype == torch.float32 else prediction.size.float()
    offset = prediction.offset if prediction.offset.dtype == torch.float32 else prediction.offset.float()
    depth = _depth_view(prediction.depth)
    B, C, H, W = hm.shape
    k = int(n_detections)
    if k > C * H * W:
        raise RuntimeError(f"selected index k out of range (k={k} > {C * H * W})")
    lib = _lib.load()
    if out is not None:
        index, label, score, yx, hw, depth_out, count = (out.index, out.label, out.score, out.yx, out.hw, out.depth,
                                                         out.count)
        if tuple(index.shape) != (B, k, 2) or (depth is not None) != (depth_out is not None):
            raise ValueError("`out` does not match this call's shapes")
    else:
        out = PackedDetections.allocate(B, k, depth is not None, dev)
        index, label, score, yx, hw, depth_out, count = (out.index, out.label, out.score, out.yx, out.hw, out.depth,
                                                         out.count)
    with torch.cuda.device(dev):
        nbytes = lib.tauv_heatmap_topk_workspace_bytes(B, C, H, W, k)
        ws = _lib.workspace(dev, nbytes)
        tail = (_lib.fptr(size), _lib.strides_arg(size, 4),
                _lib.fptr(offset), _lib.strides_arg(offset, 4),
                _lib.fptr(depth), _lib.strides_arg(depth, 3) if depth is not None else None,
                BOX_DECODE, int(model_config.downsample_ratio), int(model_config.in_h), int(model_config.in_w),
                float(score_threshold),
                _lib.i64ptr(index), _lib.i64ptr(label), _lib.fptr(score), _lib.dptr(yx), _lib.fptr(hw),
                _lib.fptr(depth_out), _lib.i32ptr(count), ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
        if stage_events is None:
            _lib.check(lib.tauv_centernet_decode(_lib.fptr(hm), B, C, H, W, k, *tail))
        else:
            e0, e1, e2 = stage_events
            e0.record()
            _lib.check(lib.tauv_heatmap_topk_stage1(_lib.fptr(hm), B, C, H, W, k, TOPK_SIGMOID_PEAK, ws.data_ptr(),
                                                    ws.numel(), _lib.stream_ptr(dev)))
            e1.record()
            _lib.check(lib.tauv_centernet_decode_stage2(B, C, H, W, k, *tail))
            e2.record()
    return out


def decode(prediction, model_config, n_detections: int, score_threshold: float) -> List[List[Detection]]:
    """Reference signature (decode.py:179-236): per-frame lists of ``Detection``."""
    return decode_packed(prediction, model_config, n_detections, score_threshold).to_lists()


# ------------------------------------------------------------------------------------------------
# decode_keypoints
# ------------------------------------------------------------------------------------------------

def _keypoint_map(object_config, n_channels: int):
    """(object label, keypoint slot) of every keypoint-heatmap channel, as the reference looks them up one by one
    (``object_config.decode_keypoint_index``, decode.py:104-106), and the largest keypoint count of any object."""
    table = np.empty((n_channels, 2), dtype=np.int32)
    for ch in range(n_channels):
        try:
            obj, slot = object_config.decode_keypoint_index(ch)
        except Exception:  # noqa: BLE001 - a channel no object owns can never be matched
            obj, slot = -1, -1
        table[ch] = (int(obj), int(slot))
    max_kp = max([len(c.keypoints) for c in object_config.configs] + [1])
    return table, max_kp


def decode_keypoints_packed(prediction, model_config, object_config, n_detections: int, keypoint_n_detections: int,
                            score_threshold: float, keypoint_score_threshold: float):
    """Device part of ``decode_keypoints``: three launches (objects: peaks + top-k + boxes; keypoints: peaks + top-k;
    the greedy association, csrc/centernet_keypoints.cu), nothing synchronises.  Returns (PackedDetections whose one
    storage buffer also holds kp_set / kp_yx / kp_score / kp_aff [B,k,max_kp,...], max_kp)."""
    dev = _lib.require_cuda(prediction.heatmap, prediction.keypoint_heatmap, prediction.keypoint_affinity,
                            prediction.size, prediction.depth)
    lib = _lib.load()
    hm = _as_heatmap(prediction.heatmap)
    kp_hm = _as_heatmap(prediction.keypoint_heatmap)
    B, C, H, W = hm.shape
    Kp = kp_hm.shape[1]
    k, kk = int(n_detections), int(keypoint_n_detections)
    if k > C * H * W or kk > Kp * H * W:
        raise RuntimeError("selected index k out of range")
    table, max_kp = _keypoint_map(object_config, Kp)
    size = prediction.size if prediction.size.dtype == torch.float32 else prediction.size.float()
    depth = _depth_view(prediction.depth)
    aff = prediction.keypoint_affinity
    aff = aff if aff.dtype == torch.float32 else aff.float()
    out = PackedDetections.allocate(B, k, depth is not None, dev, extra=(
        ("kp_yx", torch.float32, (B, k, max_kp, 2)), ("kp_score", torch.float32, (B, k, max_kp)),
        ("kp_aff", torch.float32, (B, k, max_kp, 2)), ("kp_set", torch.uint8, (B, k, max_kp))))
    x = out.extra
    with torch.cuda.device(dev):
        ws = _lib.workspace(dev, max(lib.tauv_heatmap_topk_workspace_bytes(B, C, H, W, k),
                                     lib.tauv_heatmap_topk_workspace_bytes(B, Kp, H, W, kk)))
        st = _lib.stream_ptr(dev)
        _lib.check(lib.tauv_centernet_decode(
            _lib.fptr(hm), B, C, H, W, k, _lib.fptr(size), _lib.strides_arg(size, 4), None, None,
            _lib.fptr(depth), _lib.strides_arg(depth, 3) if depth is not None else None,
            BOX_KEYPOINTS, int(model_config.downsample_ratio), int(model_config.in_h), int(model_config.in_w),
            float(score_threshold), _lib.i64ptr(out.index), _lib.i64ptr(out.label), _lib.fptr(out.score),
            _lib.dptr(out.yx), _lib.fptr(out.hw), _lib.fptr(out.depth), _lib.i32ptr(out.count), ws.data_ptr(), ws.numel(), st))
        kp_index = torch.empty((B, kk, 2), dtype=torch.int64, device=dev)
        kp_label = torch.empty((B, kk), dtype=torch.int64, device=dev)
        kp_score = torch.empty((B, kk), dtype=torch.float32, device=dev)
        _lib.check(lib.tauv_heatmap_topk(_lib.fptr(kp_hm), B, Kp, H, W, kk, TOPK_SIGMOID_PEAK, _lib.i64ptr(kp_index),
                                         _lib.i64ptr(kp_label), _lib.fptr(kp_score), ws.data_ptr(), ws.numel(), st))
        d_table = torch.from_numpy(table).to(dev, non_blocking=True)
        _lib.check(lib.tauv_centernet_keypoint_assoc(
            _lib.i64ptr(out.label), _lib.dptr(out.yx), _lib.i32ptr(out.count), B, k, _lib.i64ptr(kp_index),
            _lib.i64ptr(kp_label), _lib.fptr(kp_score), kk, _lib.fptr(aff), _lib.strides_arg(aff, 5),
            _lib.i32ptr(d_table), Kp, max_kp, int(model_config.out_h), int(model_config.out_w),
            float(keypoint_score_threshold), _lib.u8ptr(x["kp_set"]), _lib.fptr(x["kp_yx"]), _lib.fptr(x["kp_score"]),
            _lib.fptr(x["kp_aff"]), st))
    return out, max_kp


def decode_keypoints(prediction, model_config, object_config, M_projection: np.ndarray,
                     n_detections: int, keypoint_n_detections: int,
                     score_threshold: float, keypoint_score_threshold: float,
                     keypoint_angle_threshold: float) -> List[List[KeypointDetection]]:
    """Reference signature (decode.py:51-176).

    Device: both fused peak/top-k passes, the object box gather and the greedy keypoint->object association
    (``decode_keypoints_packed``); ONE device->host copy of the packed result.  Host: building the
    ``KeypointDetection`` records and the optional PnP tail (``cv2.solvePnP``, host-side in the reference too).
    ``keypoint_angle_threshold`` is accepted and unused, exactly like the reference.
    """
    packed, max_kp = decode_keypoints_packed(prediction, model_config, object_config, n_detections,
                                             keypoint_n_detections, score_threshold, keypoint_score_threshold)
    h = packed.to_host()
    detections = []
    for b in range(h["label"].shape[0]):
        sample = []
        last_match, last_rank = None, -1
        for i in range(int(h["count"][b])):
            lab = int(h["label"][b, i])
            n_kp = len(object_config.configs[lab].keypoints)
            d = KeypointDetection(
                label=lab, score=float(h["score"][b, i]),
                y=float(h["yx"][b, i, 0]), x=float(h["yx"][b, i, 1]),
                h=float(h["hw"][b, i, 0]), w=float(h["hw"][b, i, 1]),
                depth=float(h["depth"][b, i]) if h["depth"] is not None else None,
                keypoints=[None] * n_kp, keypoint_scores=[None] * n_kp, keypoint_affinities=[None] * n_kp,
                cam_t_object=None)
            for j in range(min(n_kp, max_kp)):
                if h["kp_set"][b, i, j]:
                    d.keypoints[j] = (float(h["kp_yx"][b, i, j, 0]), float(h["kp_yx"][b, i, j, 1]))
                    d.keypoint_affinities[j] = (float(h["kp_aff"][b, i, j, 0]), float(h["kp_aff"][b, i, j, 1]))
                    d.keypoint_scores[j] = float(h["kp_score"][b, i, j])
                    # the reference's PnP tail writes to the LAST matched detection (its stale `match_detection`,
                    # decode.py:172): keypoints are matched in descending score, so that is the lowest matched score
                    if last_match is None or h["kp_score"][b, i, j] < last_rank:
                        last_match, last_rank = d, h["kp_score"][b, i, j]
            sample.append(d)
        _pnp_tail(sample, last_match, model_config, object_config, M_projection)
        detections.append(sample)
    return detections


def _pnp_tail(sample, match_detection, model_config, object_config, M_projection):
    """decode.py:137-172 — host-side cv2.solvePnP for objects with >= 6 keypoints.  The reference
    stores the pose on the *last matched* detection (``match_detection``), not on ``detection``;
    that quirk is kept.  Needs cv2 (+ spatialmath for the SE3 wrapper); silently skipped without."""
    todo = [d for d in sample if sum(kp is not None for kp in d.keypoints) >= 6]
    if not todo or match_detection is None:
        return
    try:
        import cv2
    except ImportError:  # pragma: no cover
        return
    for d in todo:
        img_pts, cam_pts = [], []
        for i, kp in enumerate(d.keypoints):
            if kp is not None:
                img_pts.append([kp[1] * model_config.in_w, kp[0] * model_config.in_h])
                cam_pts.append(object_config.configs[d.label].keypoints[i])
        ok, rvec, tvec = cv2.solvePnP(np.array(cam_pts, dtype=np.float64), np.array(img_pts, dtype=np.float64),
                                      M_projection, None, flags=cv2.SOLVEPNP_ITERATIVE)
        if ok:
            rotm, _ = cv2.Rodrigues(rvec)
            try:
                from spatialmath import SE3, SO3
                match_detection.cam_t_object = SE3.Rt(SO3(rotm), tvec)
            except ImportError:
                match_detection.cam_t_object = (rotm, tvec)


# ------------------------------------------------------------------------------------------------
# angle / depth decoders (elementwise; a6 of the scope table)
# ------------------------------------------------------------------------------------------------

def angle_get_bins(bin_overlap: float):
    """((centre, min, max) of bin 0, same of bin 1)   — decode.py:282-288."""
    bin_0 = (pi / 2, -bin_overlap / 2, pi + bin_overlap / 2)
    bin_1 = (-pi / 2, -pi - bin_overlap / 2, bin_overlap / 2)
    return bin_0, bin_1


def angle_decode(predicted_bin: torch.Tensor, predicted_offset: torch.Tensor, theta_range: float,
                 bin_overlap: float) -> torch.Tensor:
    """Two-bin angle decode [B,n,4]x2 -> [B,n]   — decode.py:291-316 (no live caller in the reference)."""
    dev = _lib.require_cuda(predicted_bin, predicted_offset)
    pb, po = _lib.f32c(predicted_bin), _lib.f32c(predicted_offset)
    n = pb.numel() // 4
    out = torch.empty(pb.shape[:-1], dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().tauv_angle_decode(_lib.fptr(pb), _lib.fptr(po), n, float(theta_range),
                                                 _lib.fptr(out), _lib.stream_ptr(dev)))
    return out


def depth_decode(prediction: torch.Tensor) -> torch.Tensor:
    """1/sigmoid(d) - 1, any shape   — decode.py:319-324."""
    dev = _lib.require_cuda(prediction)
    p = _lib.f32c(prediction)
    out = torch.empty_like(p)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().tauv_depth_decode(_lib.fptr(p), p.numel(), _lib.fptr(out), _lib.stream_ptr(dev)))
    return out
