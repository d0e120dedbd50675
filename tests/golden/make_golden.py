"""Freeze golden vectors from the REAL reference (run in the build container only).

    python tests/golden/make_golden.py            # needs /root/reference/src

Imports the unmodified TAUV-Vision modules (matplotlib / spatialmath stubbed in sys.modules — both
unused on the tensor path), runs them on CPU fp32 over the seeded inputs of
tests/synth.py and stores inputs + outputs as small .npz files next to this script.  The GPU box has
no /root/reference: tests only read the .npz files.

Cases whose order the reference leaves to torch.topk / torch.sort use "separated" inputs (all scores
distinct), so the frozen order is the only possible one.
"""
from __future__ import annotations

import sys
import types
from math import pi
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, "/root/reference/src")
for _name in ("matplotlib", "matplotlib.pyplot", "spatialmath"):
    sys.modules[_name] = types.ModuleType(_name)
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.modules["spatialmath"].SE3 = sys.modules["spatialmath"].SO3 = object

from tauv_vision.centernet.model import decode as ref_decode  # noqa: E402
from tauv_vision.centernet.model import loss as ref_loss  # noqa: E402
from tauv_vision.centernet.model.centernet import Prediction  # noqa: E402
from tauv_vision.centernet.model.config import (AngleConfig, ModelConfig, ObjectConfig, ObjectConfigSet,  # noqa: E402
                                                TrainConfig)
from tauv_vision.datasets.load.pose_dataset import PoseSample  # noqa: E402
from tauv_vision.yolact.model import anchors as ref_anchors  # noqa: E402
from tauv_vision.yolact.model import boxes as ref_boxes  # noqa: E402
from tauv_vision.yolact.model import masks as ref_masks  # noqa: E402
from tauv_vision.yolact.model import nms as ref_nms  # noqa: E402
from tauv_vision.yolact.model.config import ModelConfig as YolactModelConfig  # noqa: E402

from oracle import ref_port  # noqa: E402  (only for gaussian_splat, which the reference snapshot lacks)
from tests import synth  # noqa: E402

torch.set_num_threads(1)  # reduction order independent of the machine


def save(name, **arrays):
    out = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    np.savez_compressed(HERE / f"{name}.npz", **out)
    print(f"{name}: " + ", ".join(f"{k}{list(np.shape(v))}" for k, v in out.items()))


def cn_model_config(in_hw=512, downsamples=2):
    return ModelConfig(backbone_heights=[], backbone_channels=[], in_h=in_hw, in_w=in_hw, downsamples=downsamples,
                       angle_bin_overlap=pi / 3)


def train_config(sig_h=2.0, sig_a=3.0):
    return TrainConfig(lr=0, batch_size=0, n_batches=0, n_epochs=0, heatmap_focal_loss_a=2, heatmap_focal_loss_b=4,
                       heatmap_sigma_factor=0.1, keypoint_heatmap_sigma=sig_h, keypoint_affinity_sigma=sig_a,
                       loss_lambda_keypoint_heatmap=1, loss_lambda_keypoint_affinity=1, loss_lambda_size=1,
                       loss_lambda_offset=1, loss_lambda_angle=1, loss_lambda_depth=1, n_workers=0,
                       weight_save_interval=1)


def object_configs(n_labels, kp_per_object=0):
    ang = AngleConfig(train=False, modulo=None)
    cfgs = []
    for i in range(n_labels):
        kps = [(0.1 * j, 0.0, 0.05 * i) for j in range(kp_per_object)] if kp_per_object else None
        cfgs.append(ObjectConfig(id=f"obj{i}", yaw=ang, pitch=ang, roll=ang, train_depth=True,
                                 train_keypoints=kp_per_object > 0, keypoints=kps))
    return ObjectConfigSet(cfgs)


def prediction(heatmap, size, offset, depth=None, kp_heatmap=None, kp_affinity=None):
    return Prediction(heatmap=heatmap, keypoint_heatmap=kp_heatmap, keypoint_affinity=kp_affinity, size=size,
                      offset=offset, roll_bin=None, roll_offset=None, pitch_bin=None, pitch_offset=None,
                      yaw_bin=None, yaw_offset=None, depth=depth)


def detections_to_arrays(dets, k):
    """[[Detection]] -> padded arrays + count."""
    B = len(dets)
    lab = np.full((B, k), -1, np.int64)
    sc = np.zeros((B, k), np.float32)
    yx = np.zeros((B, k, 2), np.float64)
    hw = np.zeros((B, k, 2), np.float32)
    dep = np.zeros((B, k), np.float32)
    cnt = np.zeros((B,), np.int32)
    for b, frame in enumerate(dets):
        cnt[b] = len(frame)
        for i, d in enumerate(frame):
            lab[b, i], sc[b, i] = int(d.label), float(d.score)
            yx[b, i] = (d.y, d.x)
            hw[b, i] = (d.h, d.w)
            dep[b, i] = d.depth if d.depth is not None else 0.0
    return dict(label=lab, score=sc, yx=yx, hw=hw, depth_out=dep, count=cnt)


def main():
    # ---- KAT of decode.py:327-339 -----------------------------------------------------------------
    hm = torch.cat((ref_port.gaussian_splat(512, 512, 100, 100, 50).unsqueeze(0).unsqueeze(1),
                    ref_port.gaussian_splat(512, 512, 200, 200, 50).unsqueeze(0).unsqueeze(1)), dim=1)
    sup = ref_decode.heatmap_nms(hm, 3)
    idx, lab, sc = ref_decode.heatmap_detect(sup, 100)
    assert idx[0, 0, 0] == 100 and idx[0, 0, 1] == 100
    save("kat_decode", index=idx[:, :2], label=lab[:, :2], score=sc, n_nonzero=int((sup != 0).sum()))

    # ---- heatmap_nms / heatmap_detect, small separated ------------------------------------------------
    logits = synth.separated_logits(2, 3, 24, 28, seed=11)
    sig = torch.sigmoid(logits)
    sup = ref_decode.heatmap_nms(sig, 3)
    idx, lab, sc = ref_decode.heatmap_detect(sup, 40)
    save("cn_nms_detect", logits=logits, suppressed=sup, index=idx, label=lab, score=sc)
    # plateau behaviour of heatmap_nms (ties survive) on a quantised map
    q = torch.round(synth.natural_logits(1, 2, 16, 16, seed=12, n_peaks=(2, 4)) * 2) / 2
    save("cn_nms_plateau", heatmap=q, suppressed=ref_decode.heatmap_nms(q, 3),
         suppressed5=ref_decode.heatmap_nms(q, 5))

    # ---- decode(), separated, permuted head views, depth --------------------------------------------
    mc = cn_model_config(128, 2)  # 32x32 map
    logits = synth.separated_logits(3, 4, 32, 32, seed=21)
    size, offset, depth = synth.head_views(3, 32, 32, seed=22)
    dets = ref_decode.decode(prediction(logits, size, offset, depth), mc, 30, 0.8)
    save("cn_decode", logits=logits, size=size.contiguous(), offset=offset.contiguous(), depth=depth.contiguous(),
         in_h=128, downsamples=2, k=30, thr=0.8, **detections_to_arrays(dets, 30))
    dets0 = ref_decode.decode(prediction(logits, size, offset, None), mc, 30, 0.0)
    save("cn_decode_thr0", **detections_to_arrays(dets0, 30))

    # ---- BASELINE config #1: square-detection scenario, batch 1, 1x256x256, stride 2 ------------------
    mc1 = cn_model_config(512, 1)
    g = synth.gen(31)
    cy, cx = (int(v) for v in torch.randint(50, 207, (2,), generator=g))
    s = 50 + 100 * float(torch.rand((1,), generator=g))
    splat = ref_port.gaussian_splat(256, 256, cy, cx, 0.05 * s).clamp(1e-6, 1 - 1e-6)
    logit1 = torch.log(splat / (1 - splat)).reshape(1, 1, 256, 256)
    size1, offset1, _ = synth.head_views(1, 256, 256, seed=32, with_depth=False)
    dets1 = ref_decode.decode(prediction(logit1, size1, offset1), mc1, 100, 0.5)
    save("cn_config1", cy=cy, cx=cx, sigma=0.05 * s, **detections_to_arrays(dets1, 100))

    # ---- decode_keypoints (association on host), 2 objects x 3 keypoints ------------------------------
    oc = object_configs(2, kp_per_object=3)
    mck = cn_model_config(128, 2)
    logits = synth.separated_logits(2, 2, 32, 32, seed=41)
    kp_logits = synth.separated_logits(2, 6, 32, 32, seed=42)
    g = synth.gen(43)
    kp_aff = torch.randn((2, 6, 2, 32, 32), generator=g)
    size, offset, depth = synth.head_views(2, 32, 32, seed=44)
    kd = ref_decode.decode_keypoints(prediction(logits, size, offset, depth, kp_logits, kp_aff), mck, oc,
                                     np.eye(3), 6, 20, 0.8, 0.8, 0.3)
    flat = []
    for b, frame in enumerate(kd):
        for i, d in enumerate(frame):
            row = [b, i, d.label, d.score, d.y, d.x, d.h, d.w, d.depth]
            for j in range(3):
                kp = d.keypoints[j]
                row += [1.0, kp[0], kp[1], d.keypoint_scores[j], d.keypoint_affinities[j][0],
                        d.keypoint_affinities[j][1]] if kp is not None else [0.0] * 6
            flat.append(row)
    save("cn_decode_keypoints", logits=logits, kp_logits=kp_logits, kp_aff=kp_aff, size=size.contiguous(),
         offset=offset.contiguous(), depth=depth.contiguous(), rows=np.array(flat, np.float64),
         counts=np.array([len(f) for f in kd]))

    # ---- decode_keypoints, keypoint threshold boundary: a keypoint whose score is exactly float32(0.7) with
    # keypoint_score_threshold = 0.7.  The reference compares Python floats (decode.py:100-102):
    # float(float32(0.7)) = 0.699999988... < 0.7, so the loop stops AT that keypoint (an fp32 compare would keep it).
    x_star = None
    for cand in np.nextafter(np.float32(0.8472978), np.float32(1.0)) + np.arange(-40, 40) * np.float32(5.9604645e-08):
        if torch.sigmoid(torch.tensor(np.float32(cand))).item() == float(np.float32(0.7)):
            x_star = float(np.float32(cand))
            break
    assert x_star is not None
    kp_logits2 = synth.separated_logits(2, 6, 32, 32, seed=45, lo=-6.0, hi=x_star + 9.5e-4)
    flat0 = kp_logits2[0].reshape(-1)
    order = torch.argsort(flat0, descending=True)
    flat0[order[10]] = x_star  # the 11th best keypoint peak of frame 0 (its neighbours in rank are 0.5e-4 away)
    kd2 = ref_decode.decode_keypoints(prediction(logits, size, offset, depth, kp_logits2, kp_aff), mck, oc,
                                      np.eye(3), 6, 20, 0.8, 0.7, 0.3)
    flat2 = []
    for b, frame in enumerate(kd2):
        for i, d in enumerate(frame):
            row = [b, i, d.label, d.score, d.y, d.x, d.h, d.w, d.depth]
            for j in range(3):
                kp = d.keypoints[j]
                row += [1.0, kp[0], kp[1], d.keypoint_scores[j], d.keypoint_affinities[j][0],
                        d.keypoint_affinities[j][1]] if kp is not None else [0.0] * 6
            flat2.append(row)
    save("cn_decode_keypoints_thr", kp_logits=kp_logits2, rows=np.array(flat2, np.float64),
         counts=np.array([len(f) for f in kd2]), x_star=np.float32(x_star))

    # ---- angle_decode / depth_decode -------------------------------------------------------------------
    g = synth.gen(51)
    pb, po = torch.randn((2, 9, 4), generator=g), torch.randn((2, 9, 4), generator=g)
    save("cn_angle_depth", bin=pb, offset=po, angle=ref_decode.angle_decode(pb, po, 2 * pi, pi / 3),
         angle_pi=ref_decode.angle_decode(pb, po, pi, pi / 3), depth_in=po, depth=ref_decode.depth_decode(po))

    # ---- target encode -----------------------------------------------------------------------------------
    mce = cn_model_config(96, 2)  # 24x24 map
    tc = train_config(2.0, 3.0)
    oce = object_configs(4, kp_per_object=2)  # 4 labels, 8 keypoint channels
    t = synth.pose_truth(2, 7, 4, seed=61, n_kp_inst=9, Kp=8)
    t.center[0, 0] = torch.tensor([0.999, 0.0])  # edges
    t.center[0, 1] = t.center[0, 2]             # two objects on one cell
    t.label[0, 1] = t.label[0, 2]
    t.valid[0, :3] = True
    truth = PoseSample(img=None, valid=t.valid, label=t.label, center=t.center, size=t.size, roll=None, pitch=None,
                       yaw=None, depth=None, keypoint_valid=t.keypoint_valid, keypoint_label=t.keypoint_label,
                       keypoint_center=t.keypoint_center, keypoint_object_index=t.keypoint_object_index)
    hm = ref_loss.generate_heatmap(truth, mce, tc, oce)
    khm, kw, ka = ref_loss.generate_keypoint_heatmap(truth, mce, tc, oce)
    oi = ref_loss.out_index_for_position(truth.center, mce)
    pix = truth.center * torch.Tensor((mce.in_h, mce.in_w)).unsqueeze(0).unsqueeze(1)
    off = pix - mce.downsample_ratio * (pix / mce.downsample_ratio).to(torch.long)
    save("cn_encode", valid=t.valid, label=t.label, center=t.center, kp_valid=t.keypoint_valid,
         kp_label=t.keypoint_label, kp_center=t.keypoint_center, kp_obj=t.keypoint_object_index, heatmap=hm,
         kp_heatmap=khm, kp_weight=kw, kp_affinity=ka, out_index=oi, offset=off, in_h=96, downsamples=2,
         sigma_h=2.0, sigma_a=3.0)

    # ---- heatmap focal loss on the rendered target (loss.py:233-236, 302-317), with the reference's own autograd
    # gradient of the summed loss with respect to the logits; second case: no valid object (the N == 0 branch) ------
    gf = synth.gen(67)
    fl_logits = (torch.randn((2, 4, 24, 24), generator=gf) * 1.5 - 2.2)
    fl_logits[0, int(t.label[0, 0]), 23, 0] = 9.0   # a confident hit on an object centre (clamp / saturation region)
    fl_logits[1, 0, 3, 3] = 12.0                    # a confident false positive: 1 - p is tiny
    fl_logits[1, 1, 5, 5] = -15.0                   # p below the 1e-4 clamp
    fl_logits.requires_grad_(True)
    fl = ref_loss.focal_loss(torch.sigmoid(fl_logits), hm, alpha=tc.heatmap_focal_loss_a, beta=tc.heatmap_focal_loss_b)
    fl_sum = fl.sum()
    fl_grad, = torch.autograd.grad(fl_sum, fl_logits)
    none_valid = PoseSample(img=None, valid=torch.zeros_like(t.valid), label=t.label, center=t.center, size=t.size,
                            roll=None, pitch=None, yaw=None, depth=None, keypoint_valid=t.keypoint_valid,
                            keypoint_label=t.keypoint_label, keypoint_center=t.keypoint_center,
                            keypoint_object_index=t.keypoint_object_index)
    hm0 = ref_loss.generate_heatmap(none_valid, mce, tc, oce)
    fl0 = ref_loss.focal_loss(torch.sigmoid(fl_logits), hm0, alpha=tc.heatmap_focal_loss_a, beta=tc.heatmap_focal_loss_b)
    fl0_sum = fl0.sum()
    fl0_grad, = torch.autograd.grad(fl0_sum, fl_logits)
    save("cn_focal", logits=fl_logits.detach(), valid=t.valid, label=t.label, center=t.center, in_h=96, downsamples=2,
         sigma_h=2.0, alpha=float(tc.heatmap_focal_loss_a), beta=float(tc.heatmap_focal_loss_b), loss=fl.detach(),
         loss_sum=fl_sum.detach(), n_pos=int(torch.isclose(hm, torch.ones(1)).sum()), grad=fl_grad,
         loss0_sum=fl0_sum.detach(), grad0=fl0_grad)

    # ---- YOLACT anchors ------------------------------------------------------------------------------------
    ycfg = YolactModelConfig(in_w=550, in_h=550, feature_depth=0, n_classes=0, n_prototype_masks=0,
                             n_masknet_layers_pre_upsample=0, n_masknet_layers_post_upsample=0,
                             n_prediction_head_layers=0, n_classification_layers=0, n_box_layers=0, n_mask_layers=0,
                             n_fpn_downsample_layers=0, anchor_scales=(24, 48, 96, 192, 384),
                             anchor_aspect_ratios=(1 / 2, 1, 2), box_variances=(0.1, 0.2), iou_pos_threshold=0.4,
                             iou_neg_threshold=0.3, negative_example_ratio=3, img_mean=(0, 0, 0), img_stddev=(1, 1, 1))
    small = [(7, 5), (4, 3), (2, 2), (1, 1), (1, 1)]
    anc_small = torch.cat([ref_anchors.get_anchor(i, s, ycfg) for i, s in enumerate(small)], dim=1)
    full = synth.fpn_sizes(550, 550)
    anc_full = torch.cat([ref_anchors.get_anchor(i, s, ycfg) for i, s in enumerate(full)], dim=1)
    assert anc_full.shape[1] == 19248, anc_full.shape
    sel = torch.tensor([0, 1, 68, 69, 4760, 4761, 9522, 14282, 14283, 17957, 18932, 19175, 19247])
    save("yl_anchors", small_sizes=np.array(small), small=anc_small, full_sizes=np.array(full), full_sel=sel,
         full_rows=anc_full[0, sel], full_sum=anc_full.double().sum(dim=1))

    # ---- box_decode / box_encode / iou_matrix ----------------------------------------------------------
    N = anc_small.shape[1]
    g = synth.gen(71)
    enc = torch.randn((3, N, 4), generator=g) * 0.5
    dec = ref_boxes.box_decode(enc, anc_small, ycfg)
    re_enc = ref_boxes.box_encode(dec, anc_small.expand(3, -1, -1), ycfg)
    ba = torch.cat((torch.rand((2, 9, 2), generator=g), torch.rand((2, 9, 2), generator=g) * 0.5), -1)
    bb = torch.cat((torch.rand((1, 5, 2), generator=g), torch.rand((1, 5, 2), generator=g) * 0.5), -1)
    save("yl_boxes", anchor=anc_small, enc=enc, dec=dec, re_enc=re_enc, box_a=ba, box_b=bb,
         iou_ab=ref_boxes.iou_matrix(ba, bb), iou_aa=ref_boxes.iou_matrix(ba, ba),
         corners=ref_boxes.box_to_corners(ba), xy_swap=ref_boxes.box_xy_swap(ba),
         back=ref_boxes.corners_to_box(ref_boxes.box_to_corners(ba)))

    # ---- nms (frame 0 only; separated confidences) -------------------------------------------------------
    N = 1500
    g = synth.gen(81)
    anchor = torch.cat((torch.rand((1, N, 2), generator=g), torch.rand((1, N, 2), generator=g) * 0.3 + 0.05), -1)
    cls, enc = synth.yolact_heads(2, N, 6, seed=82, anchor=anchor, separated=True)
    box = ref_boxes.box_decode(enc, anchor, ycfg)
    keep = ref_nms.nms(cls, box, 120, 0.5, 0.05)
    keep_b = ref_nms.nms(cls, box, 120, 0.3, 0.5)
    keep_1 = ref_nms.nms(cls[1:], box[1:], 120, 0.5, 0.05)
    save("yl_nms", cls=cls, enc=enc, anchor=anchor, box=box, keep=keep, keep_b=keep_b, keep_frame1=keep_1,
         top_k=120)

    # ---- assemble_mask / box_to_mask ----------------------------------------------------------------------
    proto, coeff, mbox = synth.mask_inputs(8, 20, 24, 5, seed=91)
    save("yl_mask", proto=proto, coeff=coeff, box=mbox, mask=ref_masks.assemble_mask(proto, coeff, mbox),
         mask_nobox=ref_masks.assemble_mask(proto, coeff, None), crop0=ref_boxes.box_to_mask(mbox[0], (20, 24)))

    # ---- masked depth mean: the node's consumer of assemble_mask, with the reference's own calls --------------
    # (yolact_node.py:102-103 depth image, :130-131 assemble_mask + F.interpolate, :178 nanmean)
    import torch.nn.functional as F

    def node_depth_mean(proto, coeff, box, depth_mm):
        depth = depth_mm.numpy().astype(np.uint16)
        depth = np.where(depth == 0, np.nan, depth)
        depth = depth.astype(float) / 1000
        mask = ref_masks.assemble_mask(proto, coeff, box)
        mask = F.interpolate(mask.unsqueeze(0), (depth.shape[0], depth.shape[1])).squeeze(0)
        means, counts = [], []
        for i in range(mask.shape[0]):
            mask_np = mask[i].detach().cpu().numpy()
            sel = np.where(mask_np > 0.5, depth, np.nan)
            counts.append(int(np.sum(~np.isnan(sel))))
            with np.errstate(all="ignore"):
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    means.append(float(np.nanmean(sel)))
        return np.array(means), np.array(counts)

    arrays = {}
    for tag, (P_, H_, W_, K_, exact) in {"a": (8, 20, 24, 5, False), "b": (32, 23, 31, 7, True)}.items():
        proto, coeff, mbox = (synth.mask_inputs_exact if exact else synth.mask_inputs)(P_, H_, W_, K_, seed=95)
        arrays.update({f"proto_{tag}": proto, f"coeff_{tag}": coeff, f"box_{tag}": mbox})
        for j, (hi, wi) in enumerate([(45, 70), (2 * H_, 2 * W_), (H_, W_), (13, 17), (97, 64)]):
            depth = synth.depth_image(hi, wi, seed=960 + j)
            mean, count = node_depth_mean(proto, coeff, mbox, depth)
            arrays.update({f"depth_{tag}{j}": depth.to(torch.int32), f"mean_{tag}{j}": mean, f"count_{tag}{j}": count})
        mean, count = node_depth_mean(proto, coeff, None, synth.depth_image(45, 70, seed=960))
        arrays.update({f"mean_{tag}_nobox": mean, f"count_{tag}_nobox": count})
    save("yl_mask_depth", **arrays)

    # ---- binarised upsampled masks with the callers' own lines: yolact_node.py:135 (+ `mask_np > 0.5`, :178) and
    # evaluate_batch.py:101-102; the resized fp32 values are kept as well (float16 is enough to tell how far a pixel
    # is from the threshold) so that the tests can set pixels within rounding of 0.5 aside
    arrays = {}
    for tag, (P_, H_, W_, K_, exact) in {"a": (8, 20, 24, 5, False), "b": (32, 23, 31, 7, True)}.items():
        proto, coeff, mbox = (synth.mask_inputs_exact if exact else synth.mask_inputs)(P_, H_, W_, K_, seed=97)
        arrays.update({f"proto_{tag}": proto, f"coeff_{tag}": coeff, f"box_{tag}": mbox})
        mask = ref_masks.assemble_mask(proto, coeff, mbox)
        for j, (ho, wo) in enumerate([(45, 72), (2 * H_, 2 * W_), (H_, W_), (13, 17), (97, 64)]):
            near = F.interpolate(mask.unsqueeze(0), (ho, wo)).squeeze(0)
            bil = F.interpolate(mask.unsqueeze(0), (ho, wo), mode="bilinear").squeeze(0)
            arrays.update({f"size_{tag}{j}": np.array([ho, wo]),
                           f"nearest_{tag}{j}": (near > 0.5).to(torch.uint8), f"bilinear_{tag}{j}": (bil > 0.5).to(torch.uint8),
                           f"bilinear_dist_{tag}{j}": (bil - 0.5).abs().to(torch.float16)})
    save("yl_mask_binary", **arrays)

    # ---- anchor matching: yolact/model/loss.py:16-22 + :62-66, line by line with the reference's own ops ----
    tb, tv = synth.truth_boxes(3, 6, seed=101)
    g = synth.gen(102)
    for b in range(3):  # half of the truths sit on (jittered) priors so that positives exist
        pick = torch.randint(0, anc_small.shape[1], (3,), generator=g)
        a = anc_small[0, pick]
        jit = torch.randn((3, 4), generator=g)
        tb[b, :3] = torch.cat((a[:, :2] + 0.15 * jit[:, :2] * a[:, 2:], a[:, 2:] * (1 + 0.15 * jit[:, 2:])), -1)
    tb[1, 4] = tb[1, 0]  # duplicated truth: the first one must win the argmax tie
    tv[0, :3] = True
    tv[1, 0] = tv[1, 4] = True
    tv[2] = False  # a frame without any valid truth
    iou = ref_boxes.iou_matrix(anc_small, tb)
    match_iou, match_index = torch.max(iou * tv.unsqueeze(1).float(), dim=2)
    positive = match_iou >= 0.4
    negative = match_iou <= 0.3
    targets, counts = [], []
    for b in range(3):
        tg = ref_boxes.box_encode(tb[b, match_index[b, positive[b]]].unsqueeze(0),
                                  anc_small[0, positive[b]].unsqueeze(0), ycfg).squeeze(0)
        targets.append(tg)
        counts.append(tg.shape[0])
    save("yl_match", anchor=anc_small, truth_box=tb, truth_valid=tv, match_iou=match_iou, match_index=match_index,
         positive=positive, negative=negative, targets=torch.cat(targets, 0), counts=np.array(counts))


def yolact_loss_golden():
    """The reference's own ``loss`` (yolact/model/loss.py:8-125) with its own autograd gradients, on a small case:
    3 frames (one without a valid truth), 159 priors, 7 classes, 8 prototypes of 20x24, 40x44 segmentation maps."""
    from tauv_vision.yolact.model import loss as ref_yloss
    ycfg = YolactModelConfig(in_w=550, in_h=550, feature_depth=0, n_classes=0, n_prototype_masks=0,
                             n_masknet_layers_pre_upsample=0, n_masknet_layers_post_upsample=0,
                             n_prediction_head_layers=0, n_classification_layers=0, n_box_layers=0, n_mask_layers=0,
                             n_fpn_downsample_layers=0, anchor_scales=(24, 48, 96, 192, 384),
                             anchor_aspect_ratios=(1 / 2, 1, 2), box_variances=(0.1, 0.2), iou_pos_threshold=0.4,
                             iou_neg_threshold=0.3, negative_example_ratio=3, img_mean=(0, 0, 0), img_stddev=(1, 1, 1))
    small = [(7, 5), (4, 3), (2, 2), (1, 1), (1, 1)]
    anchor = torch.cat([ref_anchors.get_anchor(i, s, ycfg) for i, s in enumerate(small)], dim=1)
    B, M, C1, K, PH, PW, SH, SW = 3, 6, 7, 8, 20, 24, 40, 44
    N = anchor.shape[1]
    tb, tv = synth.truth_boxes(B, M, seed=201)
    g = synth.gen(202)
    for b in range(B):  # half of the truths sit on (jittered) priors so that positives exist
        pick = torch.randint(0, N, (3,), generator=g)
        a = anchor[0, pick]
        jit = torch.randn((3, 4), generator=g)
        tb[b, :3] = torch.cat((a[:, :2] + 0.1 * jit[:, :2] * a[:, 2:], a[:, 2:] * (1 + 0.1 * jit[:, 2:])), -1)
    tb[1, 4] = tb[1, 0]  # a duplicated truth: painted over in the segmentation map, so truth 0's mask is empty (:93-94)
    tv[0, :3] = True
    tv[1, 0] = tv[1, 4] = True
    tv[2] = False        # a frame without any valid truth
    tcls = torch.randint(1, C1, (B, M), generator=g)
    seg = torch.full((B, SH, SW), 255, dtype=torch.int64)
    for b in range(B):
        for j in range(M):
            if not tv[b, j]:
                continue
            y, x, h, w = (float(v) for v in tb[b, j])
            y0, y1 = max(int((y - h / 2) * SH), 0), min(int((y + h / 2) * SH) + 1, SH)
            x0, x1 = max(int((x - w / 2) * SW), 0), min(int((x + w / 2) * SW) + 1, SW)
            seg[b, y0:y1, x0:x1] = j
    img_valid = torch.ones((B, SH, SW), dtype=torch.bool)
    img_valid[:, :3, :] = False
    img_valid[0, :, -5:] = False
    cls = (torch.randn((B, N, C1), generator=g) * 2).requires_grad_()
    enc = (torch.randn((B, N, 4), generator=g) * 0.8).requires_grad_()
    coeff = torch.tanh(torch.randn((B, N, K), generator=g)).requires_grad_()
    proto = torch.relu(torch.randn((B, K, PH, PW), generator=g)).requires_grad_()
    total, (lc, lb, lm) = ref_yloss.loss((cls, enc, coeff, anchor, proto), (tv, tcls, tb, seg, img_valid), ycfg)
    g_cls, g_enc = torch.autograd.grad(lc + lb, (cls, enc), retain_graph=True)
    g_coeff, g_proto = torch.autograd.grad(lm, (coeff, proto))
    iou = ref_boxes.iou_matrix(anchor, tb)
    match_iou = torch.max(iou * tv.unsqueeze(1).float(), dim=2).values
    save("yl_loss", anchor=anchor, truth_valid=tv, truth_cls=tcls, truth_box=tb, seg=seg, img_valid=img_valid,
         cls=cls.detach(), enc=enc.detach(), coeff=coeff.detach(), proto=proto.detach(), total=total.detach(),
         cls_loss=lc.detach(), box_loss=lb.detach(), mask_loss=lm.detach(), grad_cls=g_cls, grad_enc=g_enc,
         grad_coeff=g_coeff, grad_proto=g_proto, n_pos=int((match_iou >= 0.4).sum()), pos_thr=0.4, neg_thr=0.3,
         ratio=3, v0=0.1, v1=0.2)


def yolact_heads_golden():
    """The reference's own PredictionHead (random weights, no extra layers) on three FPN levels: the NCHW outputs of its
    three final convolutions (captured by forward hooks) and the tensors it returns, concatenated over the levels the way
    Yolact.forward does (model.py:43-58), with the reference's autograd gradients back at the convolution outputs."""
    from tauv_vision.yolact.model.prediction_head import PredictionHead
    cfg = YolactModelConfig(in_w=550, in_h=550, feature_depth=8, n_classes=6, n_prototype_masks=5,
                            n_masknet_layers_pre_upsample=0, n_masknet_layers_post_upsample=0,
                            n_prediction_head_layers=0, n_classification_layers=0, n_box_layers=0, n_mask_layers=0,
                            n_fpn_downsample_layers=0, anchor_scales=(24, 48, 96), anchor_aspect_ratios=(1 / 2, 1, 2),
                            box_variances=(0.1, 0.2), iou_pos_threshold=0.4, iou_neg_threshold=0.3,
                            negative_example_ratio=3, img_mean=(0, 0, 0), img_stddev=(1, 1, 1))
    torch.manual_seed(7)
    head = PredictionHead(cfg)
    captured = {"cls": [], "box": [], "coeff": []}
    for key, layer in (("cls", head._classification_layer), ("box", head._box_encoding_layer),
                       ("coeff", head._mask_coeff_layer)):
        def hook(_m, _i, out, key=key):
            out.retain_grad()
            captured[key].append(out)
        layer.register_forward_hook(hook)
    g = synth.gen(301)
    outs = [head(torch.randn((2, 8, h, w), generator=g)) for h, w in ((9, 7), (5, 4), (33, 2))]
    cls, box, coeff = (torch.cat([o[i] for o in outs], dim=1) for i in range(3))
    w_cls, w_box, w_coeff = (torch.randn(t.shape, generator=g) for t in (cls, box, coeff))
    ((cls * w_cls).sum() + (box * w_box).sum() + (coeff * w_coeff).sum()).backward()
    arrays = dict(cls=cls.detach(), box=box.detach(), coeff=coeff.detach(), w_cls=w_cls, w_box=w_box, w_coeff=w_coeff,
                  n_classes=6, n_prototype_masks=5, n_levels=3)
    for key in captured:
        for l, t_ in enumerate(captured[key]):
            arrays[f"{key}_level{l}"] = t_.detach()
            arrays[f"{key}_grad{l}"] = t_.grad
    save("yl_heads", **arrays)


def centernet_affinity_golden():
    """The keypoint-affinity term of the reference's loss (centernet/model/loss.py:244-246, before its lambda) on the
    targets of the reference's own generate_keypoint_heatmap, with its autograd gradient."""
    import torch.nn.functional as F
    mce = cn_model_config(96, 2)  # 24x24 map
    tc = train_config(2.0, 3.0)
    oce = object_configs(4, kp_per_object=2)  # 4 labels, 8 keypoint channels
    t = synth.pose_truth(2, 7, 4, seed=161, n_kp_inst=11, Kp=8)
    truth = PoseSample(img=None, valid=t.valid, label=t.label, center=t.center, size=t.size, roll=None, pitch=None,
                       yaw=None, depth=None, keypoint_valid=t.keypoint_valid, keypoint_label=t.keypoint_label,
                       keypoint_center=t.keypoint_center, keypoint_object_index=t.keypoint_object_index)
    _, keypoint_affinity_weight, keypoint_affinity = ref_loss.generate_keypoint_heatmap(truth, mce, tc, oce)
    pred = (torch.randn((2, 8, 2, 24, 24), generator=synth.gen(162)) * 0.7).requires_grad_()
    l = F.mse_loss(pred, keypoint_affinity, reduction="none")                       # loss.py:245
    l = (keypoint_affinity_weight.unsqueeze(2) * l).sum()                           # loss.py:246 without the lambda
    grad, = torch.autograd.grad(l, pred)
    save("cn_kp_affinity_loss", center=t.center, kp_valid=t.keypoint_valid, kp_label=t.keypoint_label,
         kp_center=t.keypoint_center, kp_obj=t.keypoint_object_index, pred=pred.detach(), loss=l.detach(), grad=grad,
         in_h=96, downsamples=2, sigma_h=2.0, sigma_a=3.0)


if __name__ == "__main__":
    if sys.argv[1:] == ["cn_kp_affinity_loss"]:
        centernet_affinity_golden()
    elif sys.argv[1:] == ["yl_loss"]:
        yolact_loss_golden()
    elif sys.argv[1:] == ["yl_heads"]:
        yolact_heads_golden()
    else:
        main()
        yolact_loss_golden()
        yolact_heads_golden()
        centernet_affinity_golden()
