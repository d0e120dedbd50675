"""CPU: the oracle (oracle/ref_port.py) against the golden vectors frozen from the REAL reference
(tests/golden/make_golden.py) and against the reference's two in-module KATs."""
from math import pi

import numpy as np
import torch

from oracle import ref_port as O
from tests import synth
from tests.helpers import assert_close, assert_equal, flat_index, golden, t


def test_kat_decode_main_block():
    """decode.py:327-339 — two sigma=50 splats; (100,100) on class 0 must rank first."""
    g = golden("kat_decode")
    hm = torch.cat((O.gaussian_splat(512, 512, 100, 100, 50)[None, None],
                    O.gaussian_splat(512, 512, 200, 200, 50)[None, None]), dim=1)
    sup = O.heatmap_nms(hm, 3)
    assert int((sup != 0).sum()) == int(g["n_nonzero"]) == 2
    idx, lab, sc = O.heatmap_detect(sup, 100)
    assert idx[0, 0].tolist() == [100, 100] and int(lab[0, 0]) == 0
    assert_equal(idx[:, :2], g["index"], "index")
    assert_equal(lab[:, :2], g["label"], "label")
    assert_equal(sc, g["score"], "score")
    # canonical order among the 98 zero-score fillers: ascending flat index
    assert flat_index(idx[0, 2:], lab[0, 2:], 512, 512).tolist() == list(range(98))


def test_yolact_boxes_main_block():
    """yolact boxes.py:106-117 round trips (with the config argument the stale block forgot)."""
    g = synth.gen(5)
    box, anchor = torch.rand((1, 1, 4), generator=g), torch.rand((1, 1, 4), generator=g)
    assert torch.allclose(box, O.corners_to_box(O.box_to_corners(box)))
    assert torch.allclose(box, O.box_decode(O.box_encode(box, anchor, (0.1, 0.2)), anchor, (0.1, 0.2)), atol=1e-6)


def test_nms_detect_golden():
    g = golden("cn_nms_detect")
    sup = O.heatmap_nms(torch.sigmoid(t(g["logits"])), 3)
    assert_equal(sup, g["suppressed"], "heatmap_nms")
    idx, lab, sc = O.heatmap_detect(sup, 40)
    assert_equal(idx, g["index"]), assert_equal(lab, g["label"]), assert_equal(sc, g["score"])


def test_nms_plateau_golden():
    g = golden("cn_nms_plateau")
    assert_equal(O.heatmap_nms(t(g["heatmap"]), 3), g["suppressed"])
    assert_equal(O.heatmap_nms(t(g["heatmap"]), 5), g["suppressed5"])


def _decode_inputs(g):
    B, H, W = g["size"].shape[:3]
    size = t(g["size"]).permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)  # NHWC view of NCHW storage
    offset = t(g["offset"]).permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
    depth = t(g["depth"]).permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
    return t(g["logits"]), size, offset, depth


def _check_packed(p, g, with_depth=True):
    cnt = g["count"]
    assert_equal(p.count, cnt, "count")
    for b in range(len(cnt)):
        n = int(cnt[b])
        assert_equal(p.label[b, :n], g["label"][b, :n], "label")
        assert_equal(p.score[b, :n], g["score"][b, :n], "score")
        assert_equal(p.yx[b, :n], g["yx"][b, :n], "yx (fp64, bit exact)")
        assert_equal(p.hw[b, :n], g["hw"][b, :n], "hw")
        if with_depth:
            assert_equal(p.depth[b, :n], g["depth_out"][b, :n], "depth")


def test_decode_golden():
    g = golden("cn_decode")
    logits, size, offset, depth = _decode_inputs(g)
    p = O.decode_packed(logits, size, offset, depth, 4, 128, 128, 30, 0.8)
    _check_packed(p, g)
    g0 = golden("cn_decode_thr0")
    p0 = O.decode_packed(logits, size, offset, None, 4, 128, 128, 30, 0.0)
    _check_packed(p0, g0, with_depth=False)
    assert (g0["count"] == 30).all()


def test_config1_square_detection_golden():
    """BASELINE.json configs[0]: batch 1, one Gaussian on a 1x256x256 map, stride 2, decode(.., 100, 0.5)."""
    g = golden("cn_config1")
    splat = O.gaussian_splat(256, 256, int(g["cy"]), int(g["cx"]), float(g["sigma"])).clamp(1e-6, 1 - 1e-6)
    logits = torch.log(splat / (1 - splat)).reshape(1, 1, 256, 256)
    size, offset, _ = synth.head_views(1, 256, 256, seed=32, with_depth=False)
    p = O.decode_packed(logits, size, offset, None, 2, 512, 512, 100, 0.5)
    _check_packed(p, g, with_depth=False)
    assert int(p.count[0]) >= 1 and p.index[0, 0].tolist() == [int(g["cy"]), int(g["cx"])]


def test_angle_depth_golden():
    g = golden("cn_angle_depth")
    assert_equal(O.angle_decode(t(g["bin"]), t(g["offset"]), 2 * pi, pi / 3), g["angle"])
    assert_equal(O.angle_decode(t(g["bin"]), t(g["offset"]), pi, pi / 3), g["angle_pi"])
    assert_equal(O.depth_decode(t(g["depth_in"])), g["depth"])


def test_encode_golden():
    g = golden("cn_encode")
    args = dict(out_h=24, out_w=24, in_h=96, in_w=96, downsample_ratio=4)
    hm = O.generate_heatmap(t(g["valid"]), t(g["label"]), t(g["center"]), 4, sigma=float(g["sigma_h"]), **args)
    assert_equal(hm, g["heatmap"], "generate_heatmap")
    khm, kw, ka = O.generate_keypoint_heatmap(t(g["kp_valid"]), t(g["kp_label"]), t(g["kp_center"]), t(g["kp_obj"]),
                                              t(g["center"]), 8, sigma_heatmap=float(g["sigma_h"]),
                                              sigma_affinity=float(g["sigma_a"]), **args)
    assert_equal(khm, g["kp_heatmap"]), assert_equal(kw, g["kp_weight"]), assert_equal(ka, g["kp_affinity"])
    assert_equal(O.out_index_for_position(t(g["center"]), 96, 96, 4, 24, 24), g["out_index"])
    assert_equal(O.offset_target(t(g["center"]), 96, 96, 4), g["offset"])


def test_focal_loss_golden():
    """loss.py:302-317 on the rendered target: the oracle restatement against the reference's own values, including
    its autograd gradient and the N == 0 branch."""
    g = golden("cn_focal")
    logits = t(g["logits"]).clone().requires_grad_(True)
    target = O.generate_heatmap(t(g["valid"]), t(g["label"]), t(g["center"]), 4, 24, 24, 96, 96, 4, float(g["sigma_h"]))
    loss = O.focal_loss(torch.sigmoid(logits), target, float(g["alpha"]), float(g["beta"]))
    assert_equal(loss.detach(), g["loss"], "elementwise focal loss")
    grad, = torch.autograd.grad(loss.sum(), logits)
    assert_equal(grad, g["grad"], "autograd gradient")
    total, n_pos = O.heatmap_focal_loss(t(g["logits"]), t(g["valid"]), t(g["label"]), t(g["center"]), 96, 96, 4,
                                        float(g["sigma_h"]), float(g["alpha"]), float(g["beta"]))
    assert_equal(total, g["loss_sum"]) and n_pos == int(g["n_pos"])
    total0, n0 = O.heatmap_focal_loss(t(g["logits"]), torch.zeros_like(t(g["valid"])), t(g["label"]), t(g["center"]), 96,
                                      96, 4, float(g["sigma_h"]), float(g["alpha"]), float(g["beta"]))
    assert_equal(total0, g["loss0_sum"]) and n0 == 0


def test_anchors_golden():
    g = golden("yl_anchors")
    cfg = synth.yolact_config()
    small = O.all_anchors([tuple(s) for s in g["small_sizes"]], cfg.anchor_scales, cfg.anchor_aspect_ratios, 550, 550)
    assert_equal(small, g["small"])
    sizes = synth.fpn_sizes(550, 550)
    assert [list(s) for s in sizes] == g["full_sizes"].tolist() == [[69, 69], [35, 35], [18, 18], [9, 9], [5, 5]]
    full = O.all_anchors(sizes, cfg.anchor_scales, cfg.anchor_aspect_ratios, 550, 550)
    assert full.shape == (1, 19248, 4)
    assert_equal(full[0, t(g["full_sel"])], g["full_rows"])
    assert_equal(full.double().sum(dim=1), g["full_sum"])


def test_boxes_golden():
    g = golden("yl_boxes")
    v = (0.1, 0.2)
    assert_equal(O.box_decode(t(g["enc"]), t(g["anchor"]), v), g["dec"])
    assert_equal(O.box_encode(t(g["dec"]), t(g["anchor"]).expand(3, -1, -1), v), g["re_enc"])
    assert_equal(O.iou_matrix(t(g["box_a"]), t(g["box_b"])), g["iou_ab"])
    assert_equal(O.iou_matrix(t(g["box_a"]), t(g["box_a"])), g["iou_aa"])
    assert_equal(O.box_to_corners(t(g["box_a"])), g["corners"])
    assert_equal(O.corners_to_box(O.box_to_corners(t(g["box_a"]))), g["back"])


def test_nms_golden():
    g = golden("yl_nms")
    cls, box = t(g["cls"]), t(g["box"])
    assert_equal(O.box_decode(t(g["enc"]), t(g["anchor"]), (0.1, 0.2)), g["box"])
    assert_equal(O.nms(cls, box, 120, 0.5, 0.05), g["keep"])
    assert_equal(O.nms(cls, box, 120, 0.3, 0.5), g["keep_b"])
    assert_equal(O.nms(cls[1:], box[1:], 120, 0.5, 0.05), g["keep_frame1"])


def test_mask_golden():
    g = golden("yl_mask")
    proto, coeff, box = t(g["proto"]), t(g["coeff"]), t(g["box"])
    assert_equal(O.assemble_mask(proto, coeff, box), g["mask"])
    assert_equal(O.assemble_mask(proto, coeff, None), g["mask_nobox"])
    assert_equal(O.box_to_mask(box[0], (20, 24)), g["crop0"])


def test_upsample_nearest_index_matches_interpolate():
    """The oracle's index map is the one F.interpolate(x, size) (the call at yolact_node.py:131) uses."""
    for in_size, out_size in [(276, 720), (276, 1280), (20, 45), (24, 70), (23, 46), (31, 31), (20, 13), (276, 277),
                              (138, 480), (7, 1000), (1000, 7)]:
        x = torch.arange(in_size, dtype=torch.float32).reshape(1, 1, 1, in_size)
        up = torch.nn.functional.interpolate(x, (1, out_size)).reshape(-1).to(torch.int64)
        assert_equal(O.upsample_nearest_index(out_size, in_size), up, f"{in_size} -> {out_size}")


def test_upsample_bilinear_taps_match_interpolate():
    """The oracle's taps and weights are the ones F.interpolate(x, size, mode="bilinear") (evaluate_batch.py:101)
    uses: interpolating a ramp and a one-hot row reproduces torch's output to rounding."""
    g = torch.Generator().manual_seed(5)
    for in_size, out_size in [(138, 550), (276, 720), (20, 45), (24, 72), (23, 46), (31, 31), (20, 13), (276, 277),
                              (7, 1000), (1000, 7), (1, 5)]:
        x = torch.rand((1, 1, 1, in_size), generator=g)
        up = torch.nn.functional.interpolate(x, (1, out_size), mode="bilinear").reshape(-1)
        i0, i1, w0, w1 = O.upsample_bilinear_taps(out_size, in_size)
        row = x.reshape(-1)
        assert_close(w0 * row[i0] + w1 * row[i1], up, rtol=0, atol=2e-7, what=f"{in_size} -> {out_size}")
        # the weights themselves, recovered exactly: row p of the resized identity holds tap p's weight per output
        eye = torch.eye(in_size).reshape(1, 1, in_size, in_size)
        wts = torch.nn.functional.interpolate(eye, (in_size, out_size), mode="bilinear")[0, 0]
        d = torch.arange(out_size)
        two = i1 > i0
        assert_equal(wts[i1, d][two], w1[two], f"right weights {in_size} -> {out_size}")
        assert_equal(wts[i0, d][two], w0[two], f"left weights {in_size} -> {out_size}")


def test_mask_binary_golden():
    """mask_binary against the callers' own lines frozen from the real reference (yolact_node.py:135 + :178,
    evaluate_batch.py:101-102).  Nearest is an index map: exact.  Bilinear: exact except where the interpolated value
    is within fp32 rounding of 0.5 (the golden carries the distance)."""
    g = golden("yl_mask_binary")
    for tag in "ab":
        proto, coeff, box = t(g[f"proto_{tag}"]), t(g[f"coeff_{tag}"]), t(g[f"box_{tag}"])
        for j in range(5):
            size = tuple(int(v) for v in g[f"size_{tag}{j}"])
            assert_equal(O.mask_binary(proto, coeff, box, size, "nearest"), g[f"nearest_{tag}{j}"], f"nearest {tag}{j}")
            got = O.mask_binary(proto, coeff, box, size, "bilinear").numpy()
            clear = g[f"bilinear_dist_{tag}{j}"].astype(np.float32) > 1e-5
            assert clear.mean() > 0.99
            assert_equal(got[clear], g[f"bilinear_{tag}{j}"][clear], f"bilinear {tag}{j}")


def test_mask_depth_golden():
    """masked_depth_mean against the node's own sequence of calls (frozen by make_golden.py)."""
    g = golden("yl_mask_depth")
    for tag in "ab":
        proto, coeff, box = t(g[f"proto_{tag}"]), t(g[f"coeff_{tag}"]), t(g[f"box_{tag}"])
        for j in range(5):
            mean, count = O.masked_depth_mean(proto, coeff, box, t(g[f"depth_{tag}{j}"]))
            assert_equal(count, g[f"count_{tag}{j}"], f"count {tag}{j}")
            assert_close(mean, g[f"mean_{tag}{j}"], rtol=1e-12, what=f"mean {tag}{j}")
        mean, count = O.masked_depth_mean(proto, coeff, None, t(g[f"depth_{tag}0"]))
        assert_equal(count, g[f"count_{tag}_nobox"]), assert_close(mean, g[f"mean_{tag}_nobox"], rtol=1e-12)
    assert np.isnan(g["mean_a3"]).any() or (g["count_a3"] > 0).all()  # (NaN rows compare as equal in assert_close)


def test_match_golden():
    g = golden("yl_match")
    mi, miou, pos, neg, tgt = O.match_anchors(t(g["anchor"]), t(g["truth_box"]), t(g["truth_valid"]), 0.4, 0.3,
                                              (0.1, 0.2))
    assert_equal(mi, g["match_index"]), assert_equal(miou, g["match_iou"])
    assert_equal(pos, g["positive"]), assert_equal(neg, g["negative"])
    assert_equal(tgt[pos], g["targets"], "box_encode targets of the positives")
    assert [int(p.sum()) for p in pos] == g["counts"].tolist()


def test_yolact_loss_golden():
    """yolact/model/loss.py:8-125 — the three terms and, through the oracle's own autograd graph, the reference's
    gradients with respect to the class logits, the box encodings, the mask coefficients and the prototypes."""
    g = golden("yl_loss")
    var = (float(g["v0"]), float(g["v1"]))
    cls, enc = t(g["cls"]).requires_grad_(), t(g["enc"]).requires_grad_()
    coeff, proto = t(g["coeff"]).requires_grad_(), t(g["proto"]).requires_grad_()
    anchor, tv, tb = t(g["anchor"]), t(g["truth_valid"]), t(g["truth_box"])
    cl, bl, sel = O.yolact_class_box_loss(cls, enc, anchor, tv, t(g["truth_cls"]), tb, float(g["pos_thr"]),
                                          float(g["neg_thr"]), var, int(g["ratio"]))
    ml = O.yolact_mask_loss(coeff, proto, anchor, tv, tb, t(g["seg"]), t(g["img_valid"]), float(g["pos_thr"]),
                            float(g["neg_thr"]), var)
    assert_equal(cl.detach(), g["cls_loss"]), assert_equal(bl.detach(), g["box_loss"]), assert_equal(ml.detach(), g["mask_loss"])
    assert int(sel.sum()) == (1 + int(g["ratio"])) * int(g["n_pos"])
    g_cls, g_enc = torch.autograd.grad(cl + bl, (cls, enc))
    g_coeff, g_proto = torch.autograd.grad(ml, (coeff, proto))
    assert_equal(g_cls, g["grad_cls"]), assert_equal(g_enc, g["grad_enc"])
    assert_equal(g_coeff, g["grad_coeff"]), assert_equal(g_proto, g["grad_proto"])


def test_pack_heads_golden():
    """The reference's PredictionHead outputs, concatenated over the levels like Yolact.forward, from the NCHW outputs of
    its final convolutions (yolact/model/prediction_head.py:111-140, model.py:55-58)."""
    g = golden("yl_heads")
    L = int(g["n_levels"])
    for key, C, th in (("cls", int(g["n_classes"]) + 1, False), ("box", 4, False), ("coeff", int(g["n_prototype_masks"]), True)):
        assert_equal(O.pack_head([t(g[f"{key}_level{l}"]) for l in range(L)], C, tanh=th), g[key], key)


def test_keypoint_affinity_loss_golden():
    """centernet/model/loss.py:244-246 — the reference's expression on its own targets, and its autograd gradient."""
    g = golden("cn_kp_affinity_loss")
    pred = t(g["pred"]).requires_grad_()
    ratio = 2 ** int(g["downsamples"])
    l = O.keypoint_affinity_loss(pred, t(g["kp_valid"]), t(g["kp_label"]), t(g["kp_center"]), t(g["kp_obj"]), t(g["center"]),
                                 24, 24, int(g["in_h"]), int(g["in_h"]), ratio, float(g["sigma_h"]), float(g["sigma_a"]))
    assert_equal(l.detach(), g["loss"])
    grad, = torch.autograd.grad(l, pred)
    assert_equal(grad, g["grad"])
