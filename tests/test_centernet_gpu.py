"""GPU parity: the CenterNet CUDA path (through the C ABI) against the golden vectors frozen from the real
reference, and against the CPU oracle on seeded inputs.  Tolerances are the north star's: indices, labels
and counts exact under the canonical tie order; fp32 values 1e-5 relative."""
from math import pi
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import ref_port as O
from tests import synth
from tests.helpers import (assert_close, assert_equal, assert_topk_tie_aware, flat_index, golden, t)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cn(cuda_device):
    from tauv_vision_b200.centernet.model import decode as D
    from tauv_vision_b200.centernet.model import loss as L
    return SimpleNamespace(D=D, L=L, dev=cuda_device)


def pred(heatmap, size, offset, depth=None, kp_heatmap=None, kp_affinity=None):
    return SimpleNamespace(heatmap=heatmap, keypoint_heatmap=kp_heatmap, keypoint_affinity=kp_affinity, size=size,
                           offset=offset, depth=depth, roll_bin=None, pitch_bin=None, yaw_bin=None)


def nhwc_view(a, dev):
    """[B,H,W,ch] numpy -> the permuted NCHW view Centernet.forward produces (centernet.py:81-89)."""
    return t(a).permute(0, 3, 1, 2).contiguous().to(dev).permute(0, 2, 3, 1)


# ---- the reference's own KAT ----------------------------------------------------------------------------

def test_kat_decode_main_block(cn):
    """decode.py:327-339 on the GPU: splats -> heatmap_nms(3) -> heatmap_detect(100)."""
    g = golden("kat_decode")
    hm = torch.cat((cn.L.gaussian_splat(512, 512, 100, 100, 50)[None, None],
                    cn.L.gaussian_splat(512, 512, 200, 200, 50)[None, None]), dim=1)
    assert_close(hm[0, 0], O.gaussian_splat(512, 512, 100, 100, 50), what="gaussian_splat")
    sup = cn.D.heatmap_nms(hm, 3)
    assert int((sup != 0).sum()) == 2
    idx, lab, sc = cn.D.heatmap_detect(sup, 100)
    assert idx[0, 0, 0] == 100 and idx[0, 0, 1] == 100  # the reference's assertion
    assert_equal(idx[:, :2], g["index"]), assert_equal(lab[:, :2], g["label"]), assert_equal(sc, g["score"])
    assert flat_index(idx[0, 2:].cpu(), lab[0, 2:].cpu(), 512, 512).tolist() == list(range(98))


# ---- heatmap_nms / heatmap_detect -------------------------------------------------------------------------

def test_heatmap_nms_golden(cn):
    g = golden("cn_nms_detect")
    sig = torch.sigmoid(t(g["logits"])).to(cn.dev)  # the reference's CPU sigmoid, so the comparison is exact
    assert_equal(cn.D.heatmap_nms(sig, 3), g["suppressed"])
    p = golden("cn_nms_plateau")
    assert_equal(cn.D.heatmap_nms(t(p["heatmap"]).to(cn.dev), 3), p["suppressed"], "plateaus survive")
    assert_equal(cn.D.heatmap_nms(t(p["heatmap"]).to(cn.dev), 5), p["suppressed5"], "kernel_size 5")
    assert_equal(cn.D.heatmap_nms(t(p["heatmap"]).to(cn.dev), 1), p["heatmap"], "kernel_size 1 is the identity")


def test_heatmap_detect_golden(cn):
    g = golden("cn_nms_detect")
    idx, lab, sc = cn.D.heatmap_detect(t(g["suppressed"]).to(cn.dev), 40)
    assert_equal(idx, g["index"]), assert_equal(lab, g["label"]), assert_equal(sc, g["score"])
    assert idx.dtype == torch.int64 and lab.dtype == torch.int64 and sc.dtype == torch.float32


def test_fused_peaks_golden(cn):
    """sigmoid + 3x3 suppression + top-k in one pass must give what the three reference calls give."""
    g = golden("cn_nms_detect")
    idx, lab, sc = cn.D.heatmap_peaks(t(g["logits"]).to(cn.dev), 40)
    assert_equal(idx, g["index"]), assert_equal(lab, g["label"])
    assert_close(sc, g["score"], what="score")


def test_reference_error_behaviour(cn):
    x = torch.zeros((1, 2, 4, 4), device=cn.dev)
    with pytest.raises(AssertionError):
        cn.D.heatmap_nms(x, 2)  # decode.py:243
    with pytest.raises(AssertionError):
        cn.D.heatmap_nms(x, 0)
    with pytest.raises(RuntimeError):
        cn.D.heatmap_detect(x, 33)  # torch.topk: selected index k out of range
    with pytest.raises(RuntimeError):
        cn.D.heatmap_detect(x.cpu(), 3)  # no CPU path
    idx, lab, sc = cn.D.heatmap_detect(x, 32)  # k == C*H*W is legal
    assert flat_index(idx[0].cpu(), lab[0].cpu(), 4, 4).tolist() == list(range(32))


# ---- decode ---------------------------------------------------------------------------------------------------

def _check_against_golden(p, g, with_depth):
    h = p.to_host()
    assert_equal(h["count"], g["count"], "count")
    for b in range(len(g["count"])):
        n = int(g["count"][b])
        assert_equal(h["label"][b, :n], g["label"][b, :n], "label")
        assert_close(h["score"][b, :n], g["score"][b, :n], what="score")
        assert_equal(h["yx"][b, :n], g["yx"][b, :n], "yx: fp64 arithmetic on exact fp32 gathers is bit exact")
        assert_equal(h["hw"][b, :n], g["hw"][b, :n], "hw")
        if with_depth:
            assert_close(h["depth"][b, :n], g["depth_out"][b, :n], what="depth")


def test_decode_golden(cn):
    g = golden("cn_decode")
    mc = synth.centernet_model_config(128, 128, 2)
    logits = t(g["logits"]).to(cn.dev)
    size, offset, depth = nhwc_view(g["size"], cn.dev), nhwc_view(g["offset"], cn.dev), nhwc_view(g["depth"], cn.dev)
    assert not size.is_contiguous()
    p = cn.D.decode_packed(pred(logits, size, offset, depth), mc, 30, 0.8)
    _check_against_golden(p, g, True)
    dets = cn.D.decode(pred(logits, size, offset, depth), mc, 30, 0.8)
    assert [len(f) for f in dets] == g["count"].tolist()
    d0 = dets[0][0]
    assert isinstance(d0, cn.D.Detection) and int(d0.label) == int(g["label"][0, 0]) and d0.y == g["yx"][0, 0, 0]
    g0 = golden("cn_decode_thr0")  # threshold 0.0 keeps all k entries (first PR-curve point of evaluate.py:213)
    p0 = cn.D.decode_packed(pred(logits, size, offset, None), mc, 30, 0.0)
    _check_against_golden(p0, g0, False)


def test_config1_square_detection(cn):
    """BASELINE.json configs[0]: batch 1, one Gaussian on 1x256x256, stride 2, decode(.., 100, 0.5)."""
    g = golden("cn_config1")
    mc = synth.centernet_model_config(512, 512, 1)
    splat = O.gaussian_splat(256, 256, int(g["cy"]), int(g["cx"]), float(g["sigma"])).clamp(1e-6, 1 - 1e-6)
    logits = torch.log(splat / (1 - splat)).reshape(1, 1, 256, 256).to(cn.dev)
    size, offset, _ = synth.head_views(1, 256, 256, seed=32, with_depth=False)
    size = size.permute(0, 3, 1, 2).contiguous().to(cn.dev).permute(0, 2, 3, 1)
    offset = offset.permute(0, 3, 1, 2).contiguous().to(cn.dev).permute(0, 2, 3, 1)
    p = cn.D.decode_packed(pred(logits, size, offset), mc, 100, 0.5)
    _check_against_golden(p, g, False)
    assert p.index[0, 0].tolist() == [int(g["cy"]), int(g["cx"])]


def test_decode_keypoints_golden(cn):
    g = golden("cn_decode_keypoints")
    mc = synth.centernet_model_config(128, 128, 2)
    kps = lambda i: [(0.1 * j, 0.0, 0.05 * i) for j in range(3)]
    configs = [SimpleNamespace(keypoints=kps(0)), SimpleNamespace(keypoints=kps(1))]
    oc = SimpleNamespace(configs=configs, decode_keypoint_index=lambda k: (int(k) // 3, int(k) % 3))
    p = pred(t(g["logits"]).to(cn.dev), nhwc_view(g["size"], cn.dev), nhwc_view(g["offset"], cn.dev),
             nhwc_view(g["depth"], cn.dev), t(g["kp_logits"]).to(cn.dev), t(g["kp_aff"]).to(cn.dev))
    out = cn.D.decode_keypoints(p, mc, oc, np.eye(3), 6, 20, 0.8, 0.8, 0.3)
    assert [len(f) for f in out] == g["counts"].tolist()
    rows = []
    for b, frame in enumerate(out):
        for i, d in enumerate(frame):
            row = [b, i, d.label, d.score, d.y, d.x, d.h, d.w, d.depth]
            for j in range(3):
                kp = d.keypoints[j]
                row += [1.0, kp[0], kp[1], d.keypoint_scores[j], d.keypoint_affinities[j][0],
                        d.keypoint_affinities[j][1]] if kp is not None else [0.0] * 6
            rows.append(row)
    assert_close(np.array(rows), g["rows"], what="KeypointDetection fields")


def test_decode_keypoints_threshold_is_a_double_compare(cn):
    """decode.py:100-102 compares Python floats: a keypoint whose score is exactly float32(0.7) is BELOW a threshold of
    0.7 (0.699999988 < 0.7) and ends the loop; an fp32 compare would keep it.  Golden from the real reference."""
    g0, g = golden("cn_decode_keypoints"), golden("cn_decode_keypoints_thr")
    mc = synth.centernet_model_config(128, 128, 2)
    kps = lambda i: [(0.1 * j, 0.0, 0.05 * i) for j in range(3)]
    oc = SimpleNamespace(configs=[SimpleNamespace(keypoints=kps(0)), SimpleNamespace(keypoints=kps(1))],
                         decode_keypoint_index=lambda k: (int(k) // 3, int(k) % 3))
    kp_logits = t(g["kp_logits"])
    assert float(torch.sigmoid(torch.tensor(float(g["x_star"]), dtype=torch.float32))) == float(np.float32(0.7))
    p = pred(t(g0["logits"]).to(cn.dev), nhwc_view(g0["size"], cn.dev), nhwc_view(g0["offset"], cn.dev),
             nhwc_view(g0["depth"], cn.dev), kp_logits.to(cn.dev), t(g0["kp_aff"]).to(cn.dev))
    out = cn.D.decode_keypoints(p, mc, oc, np.eye(3), 6, 20, 0.8, 0.7, 0.3)
    assert [len(f) for f in out] == g["counts"].tolist()
    rows = []
    for b, frame in enumerate(out):
        for i, d in enumerate(frame):
            row = [b, i, d.label, d.score, d.y, d.x, d.h, d.w, d.depth]
            for j in range(3):
                kp = d.keypoints[j]
                row += [1.0, kp[0], kp[1], d.keypoint_scores[j], d.keypoint_affinities[j][0],
                        d.keypoint_affinities[j][1]] if kp is not None else [0.0] * 6
            rows.append(row)
    assert_close(np.array(rows), g["rows"], what="KeypointDetection fields at the threshold boundary")


def test_decode_keypoints_is_one_transfer(cn):
    """The packed result (objects + associated keypoints) lives in ONE device buffer: one device->host copy."""
    g = golden("cn_decode_keypoints")
    mc = synth.centernet_model_config(128, 128, 2)
    oc = SimpleNamespace(configs=[SimpleNamespace(keypoints=[0] * 3), SimpleNamespace(keypoints=[0] * 3)],
                         decode_keypoint_index=lambda k: (int(k) // 3, int(k) % 3))
    p = pred(t(g["logits"]).to(cn.dev), nhwc_view(g["size"], cn.dev), nhwc_view(g["offset"], cn.dev),
             nhwc_view(g["depth"], cn.dev), t(g["kp_logits"]).to(cn.dev), t(g["kp_aff"]).to(cn.dev))
    packed, max_kp = cn.D.decode_keypoints_packed(p, mc, oc, 6, 20, 0.8, 0.8)
    assert max_kp == 3 and packed._storage is not None
    lo, hi = packed._storage.data_ptr(), packed._storage.data_ptr() + packed._storage.numel()
    for x in (packed.index, packed.label, packed.score, packed.yx, packed.hw, packed.depth, packed.count,
              *packed.extra.values()):
        assert lo <= x.data_ptr() < hi


# ---- oracle comparisons on seeded inputs --------------------------------------------------------------------

@pytest.mark.parametrize("B,C,H,W,k", [(2, 3, 13, 11, 17), (1, 1, 1, 5, 3), (3, 2, 40, 36, 64), (2, 5, 64, 128, 100),
                                       (1, 2, 300, 256, 50), (2, 1, 8, 2052, 20)])
def test_peaks_vs_oracle_shapes(cn, B, C, H, W, k):
    """Scalar path (W % 4 != 0), single rows, multi-item planes, wide rows."""
    logits = synth.separated_logits(B, C, H, W, seed=100 + H)
    oi, ol, osc = O.heatmap_detect(O.heatmap_nms(torch.sigmoid(logits), 3), k)
    idx, lab, sc = cn.D.heatmap_peaks(logits.to(cn.dev), k)
    assert_equal(idx, oi), assert_equal(lab, ol), assert_close(sc, osc, what="score")
    # RAW mode on arbitrary (negative too) values
    oi, ol, osc = O.heatmap_detect(logits, k)
    idx, lab, sc = cn.D.heatmap_detect(logits.to(cn.dev), k)
    assert_equal(idx, oi), assert_equal(lab, ol), assert_equal(sc, osc)


def test_unaligned_base_pointer(cn):
    """A heatmap whose storage is offset by one float takes the plain-load path and must still be right."""
    logits = synth.separated_logits(2, 3, 32, 32, seed=7)
    buf = torch.empty(logits.numel() + 1, device=cn.dev)
    buf[1:] = logits.flatten().to(cn.dev)
    view = buf[1:].view(2, 3, 32, 32)
    assert view.data_ptr() % 16 != 0
    oi, ol, _ = O.heatmap_detect(O.heatmap_nms(torch.sigmoid(logits), 3), 25)
    idx, lab, _ = cn.D.heatmap_peaks(view, 25)
    assert_equal(idx, oi), assert_equal(lab, ol)


def test_natural_frames_tie_aware(cn):
    """N(-2.2,1.5) background + planted bumps: near-equal scores may swap by an ulp of expf; everything else exact."""
    logits = synth.natural_logits(4, 8, 64, 64, seed=5)
    oi, ol, osc = O.heatmap_detect(O.heatmap_nms(torch.sigmoid(logits), 3), 100)
    idx, lab, sc = cn.D.heatmap_peaks(logits.to(cn.dev), 100)
    swaps = assert_topk_tie_aware(flat_index(idx.cpu(), lab.cpu(), 64, 64), sc.cpu(), flat_index(oi, ol, 64, 64), osc,
                                  what="natural", allow_swaps=2)
    assert swaps <= 2


def test_plateaus_and_fillers(cn):
    dev = cn.dev
    # constant logits: every cell is a peak with score 0.5 -> lowest flat indices win
    idx, lab, sc = cn.D.heatmap_peaks(torch.zeros((2, 3, 16, 16), device=dev), 40)
    assert flat_index(idx[1].cpu(), lab[1].cpu(), 16, 16).tolist() == list(range(40))
    assert (sc == 0.5).all()
    # saturated sigmoid: logits 20 and 30 both give exactly 1.0f -> a plateau in score space, as in the reference
    x = torch.full((1, 1, 8, 8), -5.0)
    x[0, 0, 2, 2], x[0, 0, 2, 3] = 20.0, 30.0
    oi, ol, osc = O.heatmap_detect(O.heatmap_nms(torch.sigmoid(x), 3), 4)
    idx, lab, sc = cn.D.heatmap_peaks(x.to(dev), 4)
    assert osc[0, :2].tolist() == [1.0, 1.0]
    assert_equal(idx, oi), assert_equal(sc, osc)
    # fewer positive peaks than k: zero-score fillers in ascending flat index, skipping the peak cells
    x = torch.full((1, 2, 6, 6), -200.0)  # sigmoid underflows to exactly 0
    x[0, 0, 0, 1], x[0, 1, 3, 3] = 1.0, 2.0
    oi, ol, osc = O.heatmap_detect(O.heatmap_nms(torch.sigmoid(x), 3), 10)
    idx, lab, sc = cn.D.heatmap_peaks(x.to(dev), 10)
    assert_equal(idx, oi), assert_equal(lab, ol), assert_close(sc, osc)
    assert flat_index(idx[0, 2:].cpu(), lab[0, 2:].cpu(), 6, 6).tolist() == [0, 2, 3, 4, 5, 6, 7, 8]
    # -inf everywhere
    idx, lab, sc = cn.D.heatmap_peaks(torch.full((1, 1, 4, 4), float("-inf"), device=dev), 5)
    assert (sc == 0).all() and flat_index(idx[0].cpu(), lab[0].cpu(), 4, 4).tolist() == [0, 1, 2, 3, 4]


def test_large_k_and_pruning(cn):
    """k = 1000 and a plateau-heavy map that overflows the per-item candidate list (exercises the prune path)."""
    logits = synth.separated_logits(2, 4, 128, 128, seed=9)
    oi, ol, osc = O.heatmap_detect(O.heatmap_nms(torch.sigmoid(logits), 3), 1000)
    idx, lab, sc = cn.D.heatmap_peaks(logits.to(cn.dev), 1000)
    assert_equal(idx, oi), assert_equal(lab, ol), assert_close(sc, osc)
    q = (torch.round(synth.natural_logits(1, 2, 128, 128, seed=10) * 0.5) / 0.5)  # coarse quantisation: huge plateaus
    oi, ol, osc = O.heatmap_detect(O.heatmap_nms(torch.sigmoid(q), 3), 300)
    idx, lab, sc = cn.D.heatmap_peaks(q.to(cn.dev), 300)
    assert_equal(idx, oi), assert_equal(lab, ol), assert_close(sc, osc)
    oi, ol, osc = O.heatmap_detect(q, 300)  # RAW mode: every element is a candidate
    idx, lab, sc = cn.D.heatmap_detect(q.to(cn.dev), 300)
    assert_equal(idx, oi), assert_equal(lab, ol), assert_equal(sc, osc)


def test_many_frames_several_units_per_cluster(cn):
    """More frames than resident clusters (a cluster works through several frames, the fused tail runs once per
    frame), a frame count that is not a multiple of anything, and the threshold count of decode on top."""
    B, C, H, W, k = 150, 8, 64, 64, 50
    logits = synth.separated_logits(B, C, H, W, seed=77, lo=-8.0, hi=3.0)
    size, offset, depth = synth.head_views(B, H, W, seed=78)
    o = O.decode_packed(logits, size, offset, depth, 8, 512, 512, k, 0.6)
    mc = synth.centernet_model_config(512, 512, 3)
    dv = lambda a: a.permute(0, 3, 1, 2).contiguous().to(cn.dev).permute(0, 2, 3, 1)
    p = cn.D.decode_packed(pred(logits.to(cn.dev), dv(size), dv(offset), dv(depth)), mc, k, 0.6)
    assert_equal(p.index, o.index), assert_equal(p.label, o.label), assert_close(p.score, o.score)
    assert_equal(p.yx, o.yx), assert_equal(p.hw, o.hw), assert_equal(p.count, o.count)
    oi, ol, osc = O.heatmap_detect(logits, k)  # RAW mode through the same kernel
    idx, lab, sc = cn.D.heatmap_detect(logits.to(cn.dev), k)
    assert_equal(idx, oi), assert_equal(lab, ol), assert_equal(sc, osc)


def test_decode_full_size_vs_oracle(cn):
    """BASELINE.json configs[1] shape per frame (C=80, 128x128, k=100, stride 4) on a few frames, plus the
    other two strides of the multi-scale config at reduced batch."""
    for (B, H, ds, seed) in [(4, 128, 2, 1), (2, 256, 1, 2), (4, 64, 3, 3)]:
        logits = synth.separated_logits(B, 80, H, H, seed=seed, lo=-9.0, hi=4.0)
        size, offset, depth = synth.head_views(B, H, H, seed=seed + 10)
        o = O.decode_packed(logits, size, offset, depth, 2 ** ds, 512, 512, 100, 0.3)
        mc = synth.centernet_model_config(512, 512, ds)
        dv = lambda a: a.permute(0, 3, 1, 2).contiguous().to(cn.dev).permute(0, 2, 3, 1)
        p = cn.D.decode_packed(pred(logits.to(cn.dev), dv(size), dv(offset), dv(depth)), mc, 100, 0.3)
        assert_equal(p.index, o.index), assert_equal(p.label, o.label), assert_close(p.score, o.score)
        assert_equal(p.yx, o.yx), assert_equal(p.hw, o.hw), assert_close(p.depth, o.depth), assert_equal(p.count, o.count)


def test_full_batch_properties(cn):
    """Size-independent properties at BASELINE.json's full size (B=64, C=80, 128x128, k=100): scores sorted,
    every reported cell is a 3x3 maximum of its plane, the k-th score bounds every unreported peak, and the
    result is idempotent under a second run."""
    torch.manual_seed(0)
    logits = torch.randn((64, 80, 128, 128), device=cn.dev) * 1.5 - 2.2
    idx, lab, sc = cn.D.heatmap_peaks(logits, 100)
    assert (sc[:, :-1] >= sc[:, 1:]).all()
    sig = torch.sigmoid(logits)
    pooled = torch.nn.functional.max_pool2d(sig, 3, 1, 1)
    b = torch.arange(64, device=cn.dev)[:, None].expand(-1, 100)
    assert (pooled[b, lab, idx[..., 0], idx[..., 1]] == sig[b, lab, idx[..., 0], idx[..., 1]]).all()
    assert_close(sig[b, lab, idx[..., 0], idx[..., 1]], sc, what="reported score is the cell's sigmoid")
    sup = torch.where(pooled == sig, sig, torch.zeros_like(sig)).reshape(64, -1)
    kth = torch.topk(sup, 100).values[:, -1]
    assert_close(sc[:, -1], kth, what="k-th score")
    idx2, lab2, sc2 = cn.D.heatmap_peaks(logits, 100)
    assert_equal(idx, idx2), assert_equal(lab, lab2), assert_equal(sc, sc2)


def _smooth_ranked_logits(B, C, H, W, seed, box=9, passes=1, lo=-6.0, hi=3.0):
    """A smooth random field (box-filtered noise) whose cells carry the values of separated_logits' ramp in the
    field's rank order: few 3x3 peaks, hot regions that span many 4x8 blocks, and still no near-ties at the top."""
    g = synth.gen(seed)
    ramp = synth.separated_logits(1, C, H, W, seed, lo=lo, hi=hi).flatten().sort().values
    out = torch.empty((B, C * H * W), dtype=torch.float32)
    for b in range(B):
        f = torch.randn((1, C, H + passes * (box - 1), W + passes * (box - 1)), generator=g, dtype=torch.float64)
        for _ in range(passes):
            f = torch.nn.functional.avg_pool2d(f, box, 1)
        f = f.flatten() + 1e-9 * torch.rand(f.numel(), generator=g, dtype=torch.float64)
        out[b, f.argsort()] = ramp
    return out.reshape(B, C, H, W)


@pytest.mark.parametrize("B,C,H,W,k,box,passes", [(2, 4, 64, 64, 50, 9, 1), (1, 8, 128, 128, 100, 5, 1),
                                                  (1, 2, 96, 160, 256, 15, 1), (2, 6, 128, 128, 100, 9, 3)])
def test_select_path_smooth_fields(cn, B, C, H, W, k, box, passes):
    """Block maxima that are mostly NOT peaks (smooth maps): the select pass must lower its threshold / fall back to
    the exhaustive segments and still return the reference's ranked peaks exactly."""
    logits = _smooth_ranked_logits(B, C, H, W, seed=300 + box, box=box, passes=passes)
    oi, ol, osc = O.heatmap_detect(O.heatmap_nms(torch.sigmoid(logits), 3), k)
    idx, lab, sc = cn.D.heatmap_peaks(logits.to(cn.dev), k)
    assert_equal(idx, oi), assert_equal(lab, ol), assert_close(sc, osc, what="score")


@pytest.mark.parametrize("q,hi", [(0.25, 3.0), (0.001, 3.0), (0.0005, 12.0)])
def test_select_path_smooth_plateaus_and_saturation(cn, q, hi):
    """Smooth fields quantised to steps of q: plateaus that span block borders (every plateau cell is a peak), steps
    below the margin of the in-block pre-filter of the select pass (1e-3), and logits where the fp32 sigmoid saturates
    so that cells BELOW a neighbour still tie with it — the queue filter may only drop what sel_is_peak would."""
    g = synth.gen(77)
    f = torch.randn((2, 6, 64 + 16, 64 + 16), generator=g)
    for _ in range(2):
        f = torch.nn.functional.avg_pool2d(f, 9, 1)
    f = (f - f.min()) / (f.max() - f.min()) * (hi + 6.0) - 6.0
    logits = (torch.round(f / q) * q).contiguous()
    oi, ol, osc = O.heatmap_detect(O.heatmap_nms(torch.sigmoid(logits), 3), 100)
    idx, lab, sc = cn.D.heatmap_peaks(logits.to(cn.dev), 100)
    assert_equal(idx, oi), assert_equal(lab, ol), assert_close(sc, osc, what="score")


@pytest.mark.parametrize("B,C,H,W,k", [(2, 3, 13, 12, 17), (1, 1, 1, 8, 3), (3, 2, 41, 36, 64), (1, 7, 50, 20, 256),
                                       (2, 16, 128, 128, 256), (5, 1, 128, 128, 100), (1, 80, 128, 128, 1),
                                       (2, 20, 128, 128, 1000), (1, 80, 128, 128, 500), (2, 12, 128, 128, 600),
                                       (1, 40, 64, 64, 1024)])
def test_select_path_shapes(cn, B, C, H, W, k):
    """Aligned maps with W % 4 == 0 take the block-maxima + select path: partial row groups (H % 8 != 0), rows
    narrower than a warp of blocks, one-plane frames (block-level threshold), k = 1, and the large-k variants: four
    strided maxima per thread for the threshold (K1 > 1024), candidate lists ranked by a sort instead of by counting,
    k at the path's limit of 1024."""
    logits = synth.separated_logits(B, C, H, W, seed=500 + H + k)
    oi, ol, osc = O.heatmap_detect(O.heatmap_nms(torch.sigmoid(logits), 3), k)
    idx, lab, sc = cn.D.heatmap_peaks(logits.to(cn.dev), k)
    assert_equal(idx, oi), assert_equal(lab, ol), assert_close(sc, osc, what="score")


def test_select_path_monotone_and_sparse(cn):
    """Maps with almost no peaks: a ramp along the flat index (one peak per plane: the last cell) and a map that is
    -200 except three cells — fewer than k positive peaks, so the exhaustive pass runs and the fillers follow."""
    ramp = torch.linspace(-4.0, 4.0, 2 * 32 * 32).reshape(1, 2, 32, 32)
    x = torch.full((1, 3, 24, 40), -200.0)
    x[0, 0, 5, 7], x[0, 2, 23, 39], x[0, 1, 0, 0] = 0.5, 1.5, -1.0
    for m, k in ((ramp, 20), (x, 12)):
        oi, ol, osc = O.heatmap_detect(O.heatmap_nms(torch.sigmoid(m), 3), k)
        idx, lab, sc = cn.D.heatmap_peaks(m.to(cn.dev), k)
        assert_equal(idx, oi), assert_equal(lab, ol), assert_close(sc, osc, what="score")


def test_select_path_repeatable_under_load(cn):
    """Run-to-run determinism of the two-launch path (programmatic dependent launch, shared-memory appends in
    arbitrary order): 20 back-to-back runs on plateau-heavy and natural frames give identical outputs."""
    q = torch.round(synth.natural_logits(6, 8, 64, 64, seed=11) * 2) / 2
    nat = synth.natural_logits(6, 8, 64, 64, seed=12)
    for m in (q.to(cn.dev), nat.to(cn.dev)):
        first = cn.D.heatmap_peaks(m, 100)
        for _ in range(20):
            again = cn.D.heatmap_peaks(m, 100)
            for a, b in zip(first, again):
                assert_equal(a, b)
    oi, ol, osc = O.heatmap_detect(O.heatmap_nms(torch.sigmoid(q), 3), 100)
    idx, lab, sc = cn.D.heatmap_peaks(q.to(cn.dev), 100)
    assert_equal(idx, oi), assert_equal(lab, ol), assert_close(sc, osc, what="score")


# ---- target encode -----------------------------------------------------------------------------------------------

def test_encode_golden(cn):
    g = golden("cn_encode")
    mc = synth.centernet_model_config(96, 96, 2)
    tc = SimpleNamespace(keypoint_heatmap_sigma=float(g["sigma_h"]), keypoint_affinity_sigma=float(g["sigma_a"]))
    oc = SimpleNamespace(n_labels=4, n_keypoints=8)
    dev = cn.dev
    truth = SimpleNamespace(valid=t(g["valid"]).to(dev), label=t(g["label"]).to(dev), center=t(g["center"]).to(dev),
                            keypoint_valid=t(g["kp_valid"]).to(dev), keypoint_label=t(g["kp_label"]).to(dev),
                            keypoint_center=t(g["kp_center"]).to(dev), keypoint_object_index=t(g["kp_obj"]).to(dev))
    assert_close(cn.L.generate_heatmap(truth, mc, tc, oc), g["heatmap"], what="generate_heatmap")
    hm, wt, aff = cn.L.generate_keypoint_heatmap(truth, mc, tc, oc)
    assert_close(hm, g["kp_heatmap"], what="keypoint heatmap"), assert_close(wt, g["kp_weight"], what="affinity weight")
    # ATen's vectorised CPU sqrt is not correctly rounded (off by one ulp on ~0.3 % of pixels), so the unit
    # vectors agree to an ulp rather than bit for bit
    assert_close(aff, g["kp_affinity"], atol=1e-7, what="affinity field")
    assert_equal(cn.L.out_index_for_position(truth.center, mc), g["out_index"])
    assert_equal(cn.L.offset_target(truth.center, mc), g["offset"])


@pytest.mark.parametrize("B,n,C,H,W,ds", [(3, 5, 4, 13, 11, 1), (2, 16, 80, 128, 128, 2), (1, 0, 2, 8, 8, 2),
                                              (2, 40, 19, 24, 36, 2), (1, 32, 3, 40, 20, 1), (2, 7, 9, 30, 44, 2)])
def test_encode_vs_oracle(cn, B, n, C, H, W, ds):
    ratio = 2 ** ds
    mc = SimpleNamespace(in_h=H * ratio, in_w=W * ratio, downsample_ratio=ratio, out_h=H, out_w=W)
    tc = SimpleNamespace(keypoint_heatmap_sigma=1.7, keypoint_affinity_sigma=0.05)  # 0.05: no sigma floor in the keypoint path
    tr = synth.pose_truth(B, max(n, 1), C, seed=H, n_kp_inst=max(n, 1), Kp=3)
    if n == 0:
        tr.valid[:] = False
        tr.keypoint_valid[:] = False
    oc = SimpleNamespace(n_labels=C, n_keypoints=3)
    args = dict(out_h=H, out_w=W, in_h=mc.in_h, in_w=mc.in_w, downsample_ratio=ratio)
    o = O.generate_heatmap(tr.valid, tr.label, tr.center, C, sigma=1.7, **args)
    dtr = synth.truth_to(tr, cn.dev)
    assert_close(cn.L.generate_heatmap(dtr, mc, tc, oc), o, what="generate_heatmap")
    okh, okw, oka = O.generate_keypoint_heatmap(tr.keypoint_valid, tr.keypoint_label, tr.keypoint_center,
                                                tr.keypoint_object_index, tr.center, 3, sigma_heatmap=1.7,
                                                sigma_affinity=0.05, **args)
    hm, wt, aff = cn.L.generate_keypoint_heatmap(dtr, mc, tc, oc)
    assert_close(hm, okh), assert_close(wt, okw), assert_close(aff, oka, atol=1e-7, what="affinity")


def test_encode_sigma_floor_and_far_centres(cn):
    mc = SimpleNamespace(in_h=64, in_w=64, downsample_ratio=4, out_h=16, out_w=16)
    tc = SimpleNamespace(keypoint_heatmap_sigma=0.01, keypoint_affinity_sigma=1.0)  # floored to 0.1 (loss.py:60-62)
    tr = SimpleNamespace(valid=torch.tensor([[True, True]]), label=torch.tensor([[0, 1]]),
                         center=torch.tensor([[[0.5, 0.5], [40.0, -3.0]]]))  # second centre far outside, unclamped
    o = O.generate_heatmap(tr.valid, tr.label, tr.center, 2, 16, 16, 64, 64, 4, 0.01)
    got = cn.L.generate_heatmap(synth.truth_to(tr, cn.dev), mc, tc, SimpleNamespace(n_labels=2))
    assert_close(got, o)
    assert float(got[0, 0, 8, 8]) == 1.0 and float(got[0, 1].max()) == 0.0


def test_encode_full_size_properties(cn):
    """B=64, C=80, 128x128, 16 objects: every valid object's cell is exactly 1.0, planes without objects are 0."""
    tr = synth.pose_truth(64, 16, 80, seed=3)
    mc = synth.centernet_model_config(512, 512, 2)
    hm = cn.L.generate_heatmap(synth.truth_to(tr, cn.dev), mc, SimpleNamespace(keypoint_heatmap_sigma=2.0),
                               SimpleNamespace(n_labels=80))
    cell = O.out_index_for_position(tr.center, 512, 512, 4, 128, 128)
    b = torch.arange(64)[:, None].expand(-1, 16)
    peak = hm.cpu()[b, tr.label, cell[..., 0], cell[..., 1]]
    assert (peak[tr.valid] == 1.0).all()
    has = torch.zeros((64, 80), dtype=torch.bool)
    has[b[tr.valid], tr.label[tr.valid]] = True
    assert (hm.cpu().amax(dim=(2, 3))[~has] == 0).all() and (hm.cpu().amax(dim=(2, 3))[has] == 1).all()


def test_cluster_kernel_repeatable_under_load(cn):
    """The round-1 cluster kernel (its lock-free strip queue between the streaming warps and the service warp) still
    serves raw-value top-k and k > 1024: 150 frames (several units per cluster) of plateau-heavy and of natural maps,
    15 back-to-back runs each, identical outputs run to run and equal to the oracle."""
    q = torch.round(synth.natural_logits(150, 4, 64, 64, seed=21) * 2) / 2
    nat = synth.natural_logits(150, 4, 64, 64, seed=22)
    for m, k in ((q, 60), (nat, 100)):
        md = m.to(cn.dev)
        first = cn.D.heatmap_detect(md, k)
        for _ in range(15):
            again = cn.D.heatmap_detect(md, k)
            for a, b in zip(first, again):
                assert_equal(a, b)
        oi, ol, osc = O.heatmap_detect(m, k)
        assert_equal(first[0], oi), assert_equal(first[1], ol), assert_equal(first[2], osc)
    big = synth.separated_logits(3, 6, 64, 64, seed=23)
    first = cn.D.heatmap_peaks(big.to(cn.dev), 1500)  # k > 1024: the cluster kernel with the merge launch
    for _ in range(5):
        again = cn.D.heatmap_peaks(big.to(cn.dev), 1500)
        for a, b in zip(first, again):
            assert_equal(a, b)
    oi, ol, osc = O.heatmap_detect(O.heatmap_nms(torch.sigmoid(big), 3), 1500)
    assert_equal(first[0], oi), assert_equal(first[1], ol), assert_close(first[2], osc, what="score")


# ---- heatmap focal loss fused with the target render --------------------------------------------

def _focal_cfgs(in_hw, ds, sigma, a, b):
    return (synth.centernet_model_config(in_hw, in_hw, ds),
            SimpleNamespace(keypoint_heatmap_sigma=sigma, heatmap_focal_loss_a=a, heatmap_focal_loss_b=b))


def test_focal_loss_golden(cn):
    """focal_loss(sigmoid(logits), generate_heatmap(truth)).sum() and its gradient against the values frozen from the
    reference (its own autograd), including the N == 0 branch.  Tolerances: the loss is a sum of 4.6 k fp32 terms
    (fp64 accumulation here, fp32 in the reference): 1e-5 relative; the gradient 1e-5 relative with 1e-9 absolute for
    the cells whose factors cancel."""
    g = golden("cn_focal")
    mc, tc = _focal_cfgs(96, 2, float(g["sigma_h"]), float(g["alpha"]), float(g["beta"]))
    dev = cn.dev
    for valid, loss_ref, grad_ref, n_ref in ((t(g["valid"]), g["loss_sum"], g["grad"], int(g["n_pos"])),
                                             (torch.zeros_like(t(g["valid"])), g["loss0_sum"], g["grad0"], 0)):
        truth = SimpleNamespace(valid=valid.to(dev), label=t(g["label"]).to(dev), center=t(g["center"]).to(dev))
        logits = t(g["logits"]).to(dev).requires_grad_(True)
        loss, n_pos = cn.L.heatmap_focal_loss(logits, truth, mc, tc, return_n_pos=True)
        assert int(n_pos) == n_ref and loss.dtype == torch.float32 and loss.dim() == 0
        assert_close(loss.detach(), loss_ref, rtol=1e-5, what="focal loss")
        (loss * 3.0).backward()  # (a non-unit upstream gradient)
        assert_close(logits.grad / 3.0, grad_ref, rtol=1e-5, atol=1e-9, what="gradient")


@pytest.mark.parametrize("B,n,C,H,W,ds,a,b", [(3, 5, 4, 13, 12, 1, 2.0, 4.0), (2, 16, 80, 128, 128, 2, 2.0, 4.0),
                                              (1, 0, 2, 8, 8, 2, 2.0, 4.0), (2, 32, 3, 40, 20, 1, 3.0, 2.0),
                                              (2, 7, 9, 30, 44, 2, 1.5, 2.5), (2, 40, 5, 24, 36, 2, 2.0, 4.0),
                                              (1, 6, 4, 16, 18, 1, 2.0, 4.0)])
def test_focal_loss_vs_oracle(cn, B, n, C, H, W, ds, a, b):
    """Forward and backward against the oracle (torch CPU, the reference's expressions with autograd) on other shapes:
    generic exponents (powf), the full configs[1] frame shape, no objects, the > 32 objects and W % 4 != 0 shapes that
    take the composed path."""
    ratio = 2 ** ds
    tr = synth.pose_truth(B, n, C, seed=40 + H)
    logits = synth.natural_logits(B, C, H, W, seed=41 + W).clamp(-12, 12)
    x = logits.clone().requires_grad_(True)
    target = O.generate_heatmap(tr.valid, tr.label, tr.center, C, H, W, H * ratio, W * ratio, ratio, 2.0)
    ref = O.focal_loss(torch.sigmoid(x), target, a, b).sum()
    gref, = torch.autograd.grad(ref, x)
    mc = SimpleNamespace(in_h=H * ratio, in_w=W * ratio, downsample_ratio=ratio, out_h=H, out_w=W)
    tc = SimpleNamespace(keypoint_heatmap_sigma=2.0, heatmap_focal_loss_a=a, heatmap_focal_loss_b=b)
    xd = logits.to(cn.dev).requires_grad_(True)
    loss, n_pos = cn.L.heatmap_focal_loss(xd, synth.truth_to(tr, cn.dev), mc, tc, return_n_pos=True)
    assert int(n_pos) == int(torch.isclose(target, torch.ones(1)).sum())
    assert_close(loss.detach(), ref.detach(), rtol=2e-5, what="focal loss")
    loss.backward()
    assert_close(xd.grad, gref, rtol=2e-5, atol=1e-9, what="gradient")


@pytest.mark.parametrize("n_inst", [24, 48])
def test_keypoint_heatmap_focal_loss_vs_oracle(cn, n_inst):
    """loss.py:238-240 — the keypoint-heatmap focal term against the oracle's generate_keypoint_heatmap + focal_loss with
    autograd (24 instances: the fused pass; 48: the composed path)."""
    B, n_obj, Kp, H, W, ratio = 2, 6, 10, 32, 40, 4
    tr = synth.pose_truth(B, n_obj, 3, seed=71, n_kp_inst=n_inst, Kp=Kp)
    logits = synth.natural_logits(B, Kp, H, W, seed=72).clamp(-12, 12)
    x = logits.clone().requires_grad_(True)
    target = O.generate_keypoint_heatmap(tr.keypoint_valid, tr.keypoint_label, tr.keypoint_center, tr.keypoint_object_index,
                                         tr.center, Kp, H, W, H * ratio, W * ratio, ratio, 2.0, 3.0)[0]
    ref = O.focal_loss(torch.sigmoid(x), target, 2.0, 4.0).sum()
    gref, = torch.autograd.grad(ref, x)
    mc = SimpleNamespace(in_h=H * ratio, in_w=W * ratio, downsample_ratio=ratio, out_h=H, out_w=W)
    tc = SimpleNamespace(keypoint_heatmap_sigma=2.0, heatmap_focal_loss_a=2.0, heatmap_focal_loss_b=4.0)
    xd = logits.to(cn.dev).requires_grad_(True)
    loss = cn.L.keypoint_heatmap_focal_loss(xd, synth.truth_to(tr, cn.dev), mc, tc)
    assert_close(loss.detach(), ref.detach(), rtol=2e-5, what="keypoint heatmap focal loss")
    loss.backward()
    assert_close(xd.grad, gref, rtol=2e-5, atol=1e-9, what="gradient")


def test_keypoint_affinity_loss_golden(cn):
    """loss.py:244-246 on the reference's own targets (tests/golden/make_golden.py), with its autograd gradient."""
    g = golden("cn_kp_affinity_loss")
    d = cn.dev
    mc = synth.centernet_model_config(int(g["in_h"]), int(g["in_h"]), int(g["downsamples"]))
    tc = SimpleNamespace(keypoint_heatmap_sigma=float(g["sigma_h"]), keypoint_affinity_sigma=float(g["sigma_a"]))
    truth = SimpleNamespace(center=t(g["center"]).to(d), keypoint_valid=t(g["kp_valid"]).to(d),
                            keypoint_label=t(g["kp_label"]).to(d), keypoint_center=t(g["kp_center"]).to(d),
                            keypoint_object_index=t(g["kp_obj"]).to(d))
    pred = t(g["pred"]).to(d).requires_grad_()
    l = cn.L.keypoint_affinity_loss(pred, truth, mc, tc)
    assert_close(l, g["loss"], what="keypoint affinity term")
    l.backward()
    assert_close(pred.grad, g["grad"], atol=1e-7, what="gradient")


@pytest.mark.parametrize("B,n_obj,Kp,m,H,W,ds", [(2, 5, 6, 40, 32, 48, 2), (1, 16, 80, 100, 128, 128, 2), (2, 3, 4, 9, 20, 18, 1),
                                                 (2, 4, 3, 150, 16, 16, 2)])
def test_keypoint_affinity_loss_vs_oracle(cn, B, n_obj, Kp, m, H, W, ds):
    """Against the oracle's autograd on other shapes: two and four instance chunks per lane (m > 32, m > 64), the full
    keypoint plane size, and the W % 4 != 0 / m > 128 shapes that take the composed path."""
    ratio = 2 ** ds
    tr = synth.pose_truth(B, n_obj, 3, seed=81 + m, n_kp_inst=m, Kp=Kp)
    pred = torch.randn((B, Kp, 2, H, W), generator=synth.gen(82 + H)) * 0.8
    x = pred.clone().requires_grad_()
    ref = O.keypoint_affinity_loss(x, tr.keypoint_valid, tr.keypoint_label, tr.keypoint_center, tr.keypoint_object_index,
                                   tr.center, H, W, H * ratio, W * ratio, ratio, 2.0, 3.0)
    gref, = torch.autograd.grad(ref, x)
    mc = SimpleNamespace(in_h=H * ratio, in_w=W * ratio, downsample_ratio=ratio, out_h=H, out_w=W)
    tc = SimpleNamespace(keypoint_heatmap_sigma=2.0, keypoint_affinity_sigma=3.0)
    xd = pred.to(cn.dev).requires_grad_()
    l = cn.L.keypoint_affinity_loss(xd, synth.truth_to(tr, cn.dev), mc, tc)
    assert_close(l.detach(), ref.detach(), rtol=2e-5, what="keypoint affinity term")
    (2 * l).backward()
    assert_close(xd.grad, 2 * gref, rtol=2e-5, atol=1e-7, what="gradient")


def test_gather_at_objects_matches_the_reference_loop(cn):
    """loss.py:196-227 — the per-object gathers of the size / offset / angle / depth heads through their permuted views,
    and the gradient autograd derives from the loop (two objects on one cell add up)."""
    B, H, W, n = 3, 24, 20, 7
    g = synth.gen(91)
    mc = SimpleNamespace(in_h=H * 4, in_w=W * 4, downsample_ratio=4, out_h=H, out_w=W)
    center = torch.rand((B, n, 2), generator=g)
    center[0, 1] = center[0, 0]
    center[0, 4] = center[0, 0]   # three objects on one cell
    idx_ref = O.out_index_for_position(center, mc.in_h, mc.in_w, 4, H, W)
    idx = cn.L.out_index_for_position(center.to(cn.dev), mc)
    assert_equal(idx, idx_ref)
    for C in (2, 4, 1):
        nchw = torch.randn((B, C, H, W), generator=g)
        wgt = torch.randn((B, n, C), generator=g)
        ref_in = nchw.clone().requires_grad_()
        view = ref_in.permute(0, 2, 3, 1)
        want = torch.zeros((B, n, C))
        rows = []
        for b in range(B):
            for o in range(n):
                rows.append(view[b, idx_ref[b, o, 0], idx_ref[b, o, 1]])     # loss.py:214-227
        want = torch.stack(rows).reshape(B, n, C)
        (want * wgt).sum().backward()
        dev_in = nchw.to(cn.dev).requires_grad_()
        got = cn.L.gather_at_objects(dev_in.permute(0, 2, 3, 1), idx)
        assert_equal(got, want.detach(), f"gather C={C}")
        (got * wgt.to(cn.dev)).sum().backward()
        assert_close(dev_in.grad, ref_in.grad, rtol=1e-6, atol=1e-7, what=f"gradient C={C}")
    depth = torch.randn((B, 1, H, W), generator=g)
    got = cn.L.gather_at_objects(depth.to(cn.dev).permute(0, 2, 3, 1)[..., 0], idx)
    assert got.shape == (B, n)
    assert_equal(got, depth[torch.arange(B)[:, None], 0, idx_ref[..., 0], idx_ref[..., 1]])


def test_focal_loss_full_batch_matches_composition(cn):
    """BASELINE configs[1] size (64 x 80 x 128 x 128): the fused pass equals the composition of this package's own
    generate_heatmap with the reference's elementwise expressions on the GPU, is deterministic run to run, and writes no
    target."""
    torch.manual_seed(3)
    dev = cn.dev
    logits = (torch.randn((64, 80, 128, 128), device=dev) * 1.5 - 2.2).requires_grad_(True)
    tr = synth.truth_to(synth.pose_truth(64, 16, 80, seed=5), dev)
    mc, tc = _focal_cfgs(512, 2, 2.0, 2.0, 4.0)
    loss = cn.L.heatmap_focal_loss(logits, tr, mc, tc)
    loss.backward()
    again = cn.L.heatmap_focal_loss(logits.detach(), tr, mc, tc)
    assert_equal(loss.detach(), again, "deterministic")
    x2 = logits.detach().clone().requires_grad_(True)
    target = cn.L.generate_heatmap(tr, mc, tc, SimpleNamespace(n_labels=80))
    ref = cn.L.focal_loss(torch.sigmoid(x2), target, 2.0, 4.0).sum()
    ref.backward()
    assert_close(loss.detach(), ref.detach(), rtol=2e-5, what="loss vs composition")
    assert_close(logits.grad, x2.grad, rtol=2e-5, atol=1e-10, what="gradient vs composition")


def test_angle_depth_golden(cn):
    g = golden("cn_angle_depth")
    pb, po = t(g["bin"]).to(cn.dev), t(g["offset"]).to(cn.dev)
    assert_close(cn.D.angle_decode(pb, po, 2 * pi, pi / 3), g["angle"], atol=1e-6, what="angle_decode")
    assert_close(cn.D.angle_decode(pb, po, pi, pi / 3), g["angle_pi"], atol=1e-6, what="angle_decode(pi)")
    assert_close(cn.D.depth_decode(t(g["depth_in"]).to(cn.dev)), g["depth"], what="depth_decode")
    assert cn.D.angle_get_bins(pi / 3) == O.angle_get_bins(pi / 3)
