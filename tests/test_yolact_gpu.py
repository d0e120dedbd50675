"""GPU parity: the YOLACT CUDA path (through the C ABI) against the golden vectors frozen from the real
reference and against the CPU oracle.  Keep sets / indices / matches exact; boxes and scores 1e-5 relative;
tensor-core mask logits 1e-2 absolute (north-star tolerances)."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import ref_port as O
from tests import synth
from tests.helpers import assert_close, assert_equal, golden, t

pytestmark = pytest.mark.gpu

CFG = synth.yolact_config()


@pytest.fixture(scope="module")
def yl(cuda_device):
    from tauv_vision_b200.yolact.model import anchors, boxes, loss, masks, nms
    return SimpleNamespace(anchors=anchors, boxes=boxes, loss=loss, masks=masks, nms=nms, dev=cuda_device)


def test_anchors_golden(yl):
    g = golden("yl_anchors")
    small = yl.anchors.all_anchors([tuple(s) for s in g["small_sizes"]], CFG, yl.dev)
    assert_equal(small, g["small"], "aspect-major prior layout, bit exact")
    lvl = yl.anchors.get_anchor(1, (4, 3), CFG, yl.dev)
    assert_equal(lvl, O.get_anchor(1, (4, 3), CFG.anchor_scales, CFG.anchor_aspect_ratios, 550, 550))
    full = yl.anchors.all_anchors(synth.fpn_sizes(550, 550), CFG, yl.dev)
    assert full.shape == (1, 19248, 4)
    assert_equal(full[0, t(g["full_sel"]).to(yl.dev)], g["full_rows"])
    assert_equal(full.double().sum(dim=1), g["full_sum"])
    # generated once per (sizes, configuration, device): the second call hands out the same tensor, cache=False a new one
    again = yl.anchors.all_anchors(synth.fpn_sizes(550, 550), CFG, yl.dev)
    assert again.data_ptr() == full.data_ptr()
    fresh = yl.anchors.all_anchors(synth.fpn_sizes(550, 550), CFG, yl.dev, cache=False)
    assert fresh.data_ptr() != full.data_ptr() and torch.equal(fresh, full)


def test_boxes_golden(yl):
    g = golden("yl_boxes")
    d = yl.dev
    anchor, enc = t(g["anchor"]).to(d), t(g["enc"]).to(d)
    assert_close(yl.boxes.box_decode(enc, anchor, CFG), g["dec"], what="box_decode")
    assert_close(yl.boxes.box_encode(t(g["dec"]).to(d), anchor.expand(3, -1, -1), CFG), g["re_enc"], atol=1e-6,
                 what="box_encode")
    ba, bb = t(g["box_a"]).to(d), t(g["box_b"]).to(d)
    assert_equal(yl.boxes.iou_matrix(ba, bb), g["iou_ab"], "iou_matrix: IEEE ops only, bit exact")
    assert_equal(yl.boxes.iou_matrix(ba, ba), g["iou_aa"])
    assert_equal(yl.boxes.iou_matrix(bb, ba), np.swapaxes(g["iou_ab"], 1, 2), "broadcast on the first operand")
    assert_equal(yl.boxes.box_to_corners(ba), g["corners"]), assert_equal(yl.boxes.box_xy_swap(ba), g["xy_swap"])
    assert_equal(yl.boxes.corners_to_box(yl.boxes.box_to_corners(ba)), g["back"])


def test_boxes_main_block_round_trip(yl):
    """yolact boxes.py:106-117."""
    # (seeded, and sizes bounded away from zero: the reference's own check draws an unseeded torch.rand box, and
    # iou(box, box) == 1 only holds to ~eps * centre / size — a 1e-3-sized box fails it on the reference too)
    g = torch.Generator(device="cpu").manual_seed(106)
    box = torch.cat((torch.rand((1, 1, 2), generator=g), torch.rand((1, 1, 2), generator=g) * 0.5 + 0.25), -1).to(yl.dev)
    anchor = (torch.rand((1, 1, 4), generator=g) + 0.05).to(yl.dev)
    assert torch.allclose(box, yl.boxes.corners_to_box(yl.boxes.box_to_corners(box)))
    assert torch.allclose(box, yl.boxes.box_decode(yl.boxes.box_encode(box, anchor, CFG), anchor, CFG), atol=1e-6)
    iou = yl.boxes.iou_matrix(box, box)
    assert iou.shape == (1, 1, 1) and abs(float(iou) - 1.0) < 1e-6
    z = torch.zeros((1, 1, 4), device=yl.dev)
    assert torch.isnan(yl.boxes.iou_matrix(z, z)).all()  # 0/0, as in the reference


def test_nms_golden(yl):
    g = golden("yl_nms")
    cls, box = t(g["cls"]).to(yl.dev), t(g["box"]).to(yl.dev)
    assert_equal(yl.nms.nms(cls, box, 120, 0.5, 0.05), g["keep"], "keep list (frame 0)")
    assert_equal(yl.nms.nms(cls, box, 120, 0.3, 0.5), g["keep_b"], "other thresholds")
    assert_equal(yl.nms.nms(cls[1:], box[1:], 120, 0.5, 0.05), g["keep_frame1"])
    keep, n_keep = yl.nms.nms_batched(cls, box, 120, 0.5, 0.05)
    assert n_keep.tolist() == [len(g["keep"]), len(g["keep_frame1"])]
    assert_equal(keep[1, : int(n_keep[1])], g["keep_frame1"], "batched entry, frame 1")
    assert yl.nms.nms(cls, box, 120, 0.5, 0.05).dtype == torch.int64


def test_detect_fused_golden(yl):
    """Decode only the ranked priors, NMS, class argmax: same keep sets as box_decode + nms of the reference."""
    g = golden("yl_nms")
    d = yl.dev
    cls, enc, anchor = t(g["cls"]).to(d), t(g["enc"]).to(d), t(g["anchor"]).to(d)
    det = yl.nms.detect(cls, enc, anchor, CFG, 120, 0.5, 0.05)
    n0, n1 = int(det.n_keep[0]), int(det.n_keep[1])
    assert_equal(det.keep[0, :n0], g["keep"]), assert_equal(det.keep[1, :n1], g["keep_frame1"])
    assert_close(det.box[0, :n0], g["box"][0][g["keep"]], what="kept boxes")
    ocls = t(g["cls"])
    assert_equal(det.class_id[0, :n0], torch.argmax(ocls[0, t(g["keep"])], dim=-1).to(torch.int32), "class ids")
    assert_close(det.score[0, :n0], O.nms_scores(ocls)[0, t(g["keep"])], what="confidences")


def test_scores_vs_oracle(yl):
    for C1 in (2, 6, 33, 81, 130):
        g = synth.gen(C1)
        cls = torch.randn((2, 700, C1), generator=g) * 3
        s, a = yl.nms.max_foreground_confidence(cls.to(yl.dev), with_argmax=True)
        assert_close(s, O.nms_scores(cls), what=f"scores C1={C1}")
        assert_equal(a, torch.argmax(cls, dim=-1).to(torch.int32), f"argmax C1={C1}")


def test_nms_edge_cases(yl):
    d = yl.dev
    g = synth.gen(3)
    # fewer priors than top_k
    cls = torch.randn((1, 7, 4), generator=g)
    box = torch.cat((torch.rand((1, 7, 2), generator=g), torch.rand((1, 7, 2), generator=g) * 0.2 + 0.05), -1)
    assert_equal(yl.nms.nms(cls.to(d), box.to(d), 50, 0.5, 0.0), O.nms(cls, box, 50, 0.5, 0.0))
    # nothing passes the confidence threshold -> empty LongTensor (yolact_node.py:131-133 checks len == 0)
    out = yl.nms.nms(cls.to(d), box.to(d), 50, 0.5, 2.0)
    assert out.shape == (0,) and out.dtype == torch.int64
    # identical boxes: only the most confident survives; equal confidences tie-break by prior index
    cls = torch.zeros((1, 40, 3))
    cls[0, :, 1] = 5.0
    box = torch.tensor([0.5, 0.5, 0.2, 0.2]).repeat(1, 40, 1)
    assert yl.nms.nms(cls.to(d), box.to(d), 40, 0.5, 0.1).tolist() == [0]
    # low-confidence and already-suppressed boxes still suppress (Fast NMS), top_k = 1000
    anchor = torch.cat((torch.rand((1, 3000, 2), generator=g), torch.rand((1, 3000, 2), generator=g) * 0.3 + 0.05), -1)
    cls, enc = synth.yolact_heads(1, 3000, 9, seed=4, anchor=anchor, n_clusters=40, per_cluster=20, separated=True)
    box = O.box_decode(enc, anchor, CFG.box_variances)
    assert_equal(yl.nms.nms(cls.to(d), box.to(d), 1000, 0.4, 0.3), O.nms(cls, box, 1000, 0.4, 0.3), "top_k 1000")


@pytest.mark.parametrize("N,top_k,plateau", [(3000, 100, False), (3000, 100, True), (1024, 512, False), (1023, 64, False),
                                             (20480, 200, False), (20481, 200, False), (5000, 513, False), (2500, 37, True)])
def test_nms_ranking_paths(yl, N, top_k, plateau):
    """The one-pass ranking (1024 <= N <= 20480, top_k <= 512) and the general radix select must give the same keep set:
    sizes on both sides of every limit, and a plateau of equal confidences (more candidates than the one-pass list
    holds: it must fall back, and equal confidences tie-break by prior index)."""
    d = yl.dev
    g = synth.gen(N + top_k)
    anchor = torch.cat((torch.rand((1, N, 2), generator=g), torch.rand((1, N, 2), generator=g) * 0.2 + 0.03), -1)
    cls, enc = synth.yolact_heads(1, N, 5, seed=N, anchor=anchor, n_clusters=30, per_cluster=10, separated=True)
    if plateau:
        cls[0, : N // 2] = 0.0  # half of the priors share one, top, confidence exactly
        cls[0, : N // 2, 1] = 9.0
    box = O.box_decode(enc, anchor, CFG.box_variances)
    assert_equal(yl.nms.nms(cls.to(d), box.to(d), top_k, 0.5, 0.05), O.nms(cls, box, top_k, 0.5, 0.05), f"N={N} top_k={top_k}")


def test_nms_threshold_on_the_boundary(yl):
    """The kernel decides `fl(inter / union) <= thr` without dividing (exact comparison against the rounding midpoint in
    double).  Thresholds that ARE one of the pair IoUs, and their float neighbours, hit the equality / midpoint cases; the
    keep set must equal the oracle's (IEEE division) every time.  Also boxes with NaN / infinite / zero sizes."""
    d = yl.dev
    g = synth.gen(77)
    N = 96
    box = torch.cat((torch.rand((1, N, 2), generator=g) * 0.5 + 0.25, torch.rand((1, N, 2), generator=g) * 0.3 + 0.1), -1)
    cls = torch.zeros((1, N, 3))
    cls[0, :, 1] = torch.linspace(6.0, 1.0, N)  # distinct, descending confidences: rank == prior index
    iou = O.iou_matrix(box, box)[0]
    vals = iou[torch.triu(torch.ones(N, N, dtype=torch.bool), 1)]
    vals = vals[(vals > 0.05) & (vals < 0.95)]
    picks = vals[torch.randperm(vals.numel(), generator=g)[:24]]
    checked = 0
    for v in picks.tolist():
        f = np.float32(v)
        for thr in (f, np.nextafter(f, np.float32(1)), np.nextafter(f, np.float32(0))):
            assert_equal(yl.nms.nms(cls.to(d), box.to(d), N, float(thr), 0.0), O.nms(cls, box, N, float(thr), 0.0),
                         f"iou_threshold {float(thr)!r}")
            checked += 1
    assert checked == 72
    weird = box.clone()
    weird[0, 3, 2] = float("nan")
    weird[0, 7, 3] = float("inf")
    weird[0, 11, 2:] = 0.0
    weird[0, 12] = weird[0, 11]           # two empty boxes at the same place: 0/0 -> NaN suppresses
    weird[0, 20, 0] = float("nan")
    assert_equal(yl.nms.nms(cls.to(d), weird.to(d), N, 0.5, 0.0), O.nms(cls, weird, N, 0.5, 0.0), "NaN / inf / empty boxes")


def test_full_size_vs_oracle(yl):
    """BASELINE.json configs[2] geometry: 19 248 priors (550x550, 3 aspect ratios), 81 classes, top_k 200."""
    d = yl.dev
    anchor = O.all_anchors(synth.fpn_sizes(550, 550), CFG.anchor_scales, CFG.anchor_aspect_ratios, 550, 550)
    cls, enc = synth.yolact_heads(3, 19248, 81, seed=6, anchor=anchor, n_clusters=15, per_cluster=12, separated=True)
    box = O.box_decode(enc, anchor, CFG.box_variances)
    scores = O.nms_scores(cls)
    det = yl.nms.detect(cls.to(d), enc.to(d), anchor.to(d), CFG, 200, 0.5, 0.05)
    for b in range(3):
        keep, ranked, conf = O.nms_frame(scores[b], box[b], 200, 0.5, 0.05)
        n = int(det.n_keep[b])
        assert_equal(det.keep[b, :n], keep, f"frame {b} keep set")
        assert 20 < n < 200
    assert_close(yl.boxes.box_decode(enc.to(d), anchor.to(d), CFG), box, what="box_decode at N=19248")


def _mask_check(yl, logits_atol):
    g = golden("yl_mask")
    d = yl.dev
    proto, coeff, box = t(g["proto"]).to(d), t(g["coeff"]).to(d), t(g["box"]).to(d)
    ref_logits = O.mask_logits(t(g["proto"]), t(g["coeff"]))
    m, lg = yl.masks.assemble_mask(proto, coeff, box, return_logits=True)
    assert_close(lg, ref_logits, rtol=0, atol=logits_atol, what="mask logits")
    assert_close(m, g["mask"], rtol=0, atol=max(logits_atol / 4, 1e-6), what="assemble_mask with crop")
    assert_equal(m == 0, g["mask"] == 0, "crop region (inclusive integer-pixel bounds) is exact")
    assert_close(yl.masks.assemble_mask(proto, coeff, None), g["mask_nobox"], rtol=0,
                 atol=max(logits_atol / 4, 1e-6), what="assemble_mask without box")
    assert_equal(yl.boxes.box_to_mask(box[0], (20, 24)), g["crop0"], "box_to_mask")
    assert yl.masks.assemble_mask(proto, coeff[:0], box[:0]).shape == (0, 20, 24)


def _off16(tn):
    """A copy of `tn` that starts 4 bytes off a 16-byte boundary.  The tensor-core kernel loads coefficient rows with
    16-byte accesses, so the dispatcher (csrc/yolact_mask_umma.cuh, umma_shape_ok) hands such a call to the CUDA-core
    kernel: the way to reach that kernel at P == 32 through the public API (release builds have no environment knobs)."""
    buf = torch.empty(tn.numel() + 4, dtype=tn.dtype, device=tn.device)
    out = buf[1:1 + tn.numel()].view(tn.shape)
    assert out.data_ptr() % 16 == 4
    out.copy_(tn)
    return out


def test_mask_golden_simt(yl):
    """P = 8 is not a tensor-core shape: the CUDA-core kernel, which keeps the reference's fp32 summation order."""
    _mask_check(yl, 2e-6)


def test_mask_golden_tensor_core(yl):
    """P = 8 is not a tensor-core shape (the golden is tiny): this exercises the dispatcher; the tensor-core
    kernel itself is checked in test_mask_vs_oracle_shapes / test_mask_logits_tensor_core."""
    _mask_check(yl, 1e-2)


def test_mask_logits_tensor_core(yl):
    """North-star tolerance: |logit error| <= 1e-2 for the bf16 tensor-core contraction.  With the hi/lo operand
    split the kernel is in fact ~1e-5 accurate; assert the spec bound and report the achieved one."""
    proto, coeff, box = synth.mask_inputs(32, 69, 69, 150, seed=77)
    ref = (coeff.double() @ proto.reshape(32, -1).double()).reshape(150, 69, 69)
    m, lg = yl.masks.assemble_mask(proto.to(yl.dev), coeff.to(yl.dev), box.to(yl.dev), return_logits=True)
    err = (lg.cpu().double() - ref).abs().max().item()
    assert err <= 1e-2, err
    assert err <= 1e-4, f"hi/lo split should give fp32-class logits, got {err}"
    m2, lg2 = yl.masks.assemble_mask(proto.to(yl.dev), _off16(coeff.to(yl.dev)), box.to(yl.dev), return_logits=True)
    assert_close(lg2, O.mask_logits(proto, coeff), rtol=0, atol=1e-5, what="CUDA-core kernel (unaligned coefficients)")
    assert not torch.equal(lg, lg2), "the two calls were meant to take different kernels"
    assert_close(m, m2, rtol=0, atol=5e-4, what="tensor-core vs CUDA-core kernel")


@pytest.mark.parametrize("P,H,W,K", [(32, 138, 138, 100), (32, 276, 276, 37), (16, 64, 40, 130), (48, 30, 36, 5),
                                      (32, 7, 9, 3), (24, 16, 16, 4)])
def test_mask_vs_oracle_shapes(yl, P, H, W, K):
    """Tensor-core path where the shape fits (P % 16 == 0, H*W % 4 == 0), CUDA-core path otherwise; both against
    the fp32 oracle.  n > 128 exercises several M tiles; odd sizes exercise ragged pixel tiles."""
    proto, coeff, box = synth.mask_inputs(P, H, W, K, seed=P + H)
    ref = torch.sigmoid(proto.reshape(P, -1).T @ coeff.T).T.reshape(K, H, W)
    crop = torch.stack([O.box_to_mask(box[i], (H, W)) for i in range(K)])
    d = yl.dev
    m = yl.masks.assemble_mask(proto.to(d), coeff.to(d), box.to(d))
    assert_close(m, ref * crop, rtol=0, atol=2.5e-3, what="mask")
    assert_equal(m == 0, (ref * crop) == 0, "crop exact")
    m2 = yl.masks.assemble_mask(proto.to(d), coeff.to(d), None)
    assert_close(m2, ref, rtol=0, atol=2.5e-3, what="mask, no crop")


def test_mask_batched_matches_single(yl):
    d = yl.dev
    B, N, P, H, W, top_k = 3, 500, 32, 48, 52, 40
    g = synth.gen(12)
    anchor = torch.cat((torch.rand((1, N, 2), generator=g), torch.rand((1, N, 2), generator=g) * 0.3 + 0.05), -1)
    cls, enc = synth.yolact_heads(B, N, 7, seed=13, anchor=anchor, separated=True)
    proto = torch.nn.functional.leaky_relu(torch.randn((B, P, H, W), generator=g)).to(d)
    coeff = torch.tanh(torch.randn((B, N, P), generator=g)).to(d)
    det = yl.nms.detect(cls.to(d), enc.to(d), anchor.to(d), CFG, top_k, 0.5, 0.05)
    out = yl.masks.assemble_mask_batched(proto, coeff, det)
    for b in range(B):
        n = int(det.n_keep[b])
        assert n > 0
        single = yl.masks.assemble_mask(proto[b], coeff[b, det.keep[b, :n]], det.box[b, :n])
        assert_equal(out[b, :n], single, f"frame {b}: batched == per-frame call")


def _u16(depth_i32, dev):
    return depth_i32.to(torch.int32).to(dev).to(torch.uint16)


@pytest.mark.parametrize("simt", [False, True])
def test_mask_depth_golden(yl, simt):
    """masked_depth_mean against the node's own call sequence (assemble_mask -> F.interpolate -> nanmean of the
    selected depth readings, yolact_node.py:102-103,130-131,178) frozen from the real reference.  Case b has
    half-integer logits from bf16-exact operands and P = 32: the tensor-core epilogue must select exactly the same
    pixels; case a (P = 8) goes through the CUDA-core kernel either way."""
    g = golden("yl_mask_depth")
    d = yl.dev
    for tag in "ab":
        proto, coeff, box = t(g[f"proto_{tag}"]).to(d), t(g[f"coeff_{tag}"]).to(d), t(g[f"box_{tag}"]).to(d)
        if simt:
            coeff = _off16(coeff)
        for j in range(5):
            mean, count = yl.masks.masked_depth_mean(proto, coeff, box, _u16(t(g[f"depth_{tag}{j}"]), d), return_count=True)
            assert_equal(count, g[f"count_{tag}{j}"], f"count {tag}{j}")
            assert_close(mean, g[f"mean_{tag}{j}"], rtol=1e-12, what=f"mean {tag}{j}")
        mean, count = yl.masks.masked_depth_mean(proto, coeff, None, _u16(t(g[f"depth_{tag}0"]), d), return_count=True)
        assert_equal(count, g[f"count_{tag}_nobox"]), assert_close(mean, g[f"mean_{tag}_nobox"], rtol=1e-12)
    e = yl.masks.masked_depth_mean(proto, coeff[:0], box[:0], _u16(t(g["depth_b0"]), d))
    assert e.shape == (0,) and e.dtype == torch.float64
    with pytest.raises(TypeError):
        yl.masks.masked_depth_mean(proto, coeff, box, t(g["depth_b0"]).to(d))  # int32 is not a mono16 image


@pytest.mark.parametrize("H,W,K,hi,wi", [(138, 138, 100, 360, 640), (276, 276, 37, 720, 1280), (69, 69, 150, 69, 69),
                                          (276, 276, 300, 240, 320)])
def test_mask_depth_vs_oracle(yl, H, W, K, hi, wi):
    """Full-size prototype maps, several M tiles (K > 256 takes two launches), up- and down-sampling.  Real-valued
    operands: a camera pixel may legitimately flip where its logit is within the bf16x2 contraction error of zero,
    so counts may differ by the number of such readings (computed in float64) and the mean accordingly."""
    proto, coeff, box = synth.mask_inputs(32, H, W, K, seed=H + K)
    depth = synth.depth_image(hi, wi, seed=hi)
    ref_mean, ref_count = O.masked_depth_mean(proto, coeff, box, depth)
    d = yl.dev
    mean, count = yl.masks.masked_depth_mean(proto.to(d), coeff.to(d), box.to(d), _u16(depth, d), return_count=True)
    lg = (coeff.double() @ proto.reshape(32, -1).double()).reshape(K, H, W)
    iy, ix = O.upsample_nearest_index(hi, H), O.upsample_nearest_index(wi, W)
    amb = ((lg.abs() < 1e-4)[:, iy][:, :, ix] & (depth != 0).unsqueeze(0)).sum(dim=(1, 2))
    diff = (count.cpu() - ref_count).abs()
    assert bool((diff <= amb).all()), f"count differs beyond the ambiguous readings: {diff.max().item()} vs {amb.max().item()}"
    clean = (amb == 0).numpy()
    assert_close(mean.cpu().numpy()[clean], ref_mean.numpy()[clean], rtol=1e-12, what="mean depth")
    assert_close(mean.cpu().numpy()[~clean], ref_mean.numpy()[~clean], rtol=1e-2, what="mean depth (ambiguous readings)")


def test_mask_depth_batched_matches_single(yl):
    d = yl.dev
    B, N, P, H, W, top_k, hi, wi = 3, 500, 32, 48, 52, 40, 120, 160
    g = synth.gen(12)
    anchor = torch.cat((torch.rand((1, N, 2), generator=g), torch.rand((1, N, 2), generator=g) * 0.3 + 0.05), -1)
    cls, enc = synth.yolact_heads(B, N, 7, seed=13, anchor=anchor, separated=True)
    proto = torch.nn.functional.leaky_relu(torch.randn((B, P, H, W), generator=g)).to(d)
    coeff = torch.tanh(torch.randn((B, N, P), generator=g)).to(d)
    depth = torch.stack([_u16(synth.depth_image(hi, wi, seed=70 + b), d) for b in range(B)])
    det = yl.nms.detect(cls.to(d), enc.to(d), anchor.to(d), CFG, top_k, 0.5, 0.05)
    mean, count = yl.masks.masked_depth_mean_batched(proto, coeff, det, depth)
    for b in range(B):
        n = int(det.n_keep[b])
        assert n > 0
        m1, c1 = yl.masks.masked_depth_mean(proto[b], coeff[b, det.keep[b, :n]], det.box[b, :n], depth[b], return_count=True)
        assert_equal(count[b, :n], c1, f"frame {b}: count")
        assert_close(mean[b, :n], m1, rtol=0, atol=0, what=f"frame {b}: batched == per-frame call")
        assert bool(torch.isnan(mean[b, n:]).all()) and int(count[b, n:].abs().sum()) == 0


@pytest.mark.parametrize("simt", [False, True])
def test_mask_binary_golden(yl, simt):
    """assemble_mask_binary against the callers' own lines frozen from the real reference: F.interpolate(mask, size)
    then `mask_np > 0.5` (yolact_node.py:135, :178) and F.interpolate(..., mode="bilinear") then `mask > 0.5`
    (evaluate_batch.py:101-102).  Case b (P = 32: the tensor-core kernel unless `simt`) has half-integer logits from
    bf16-exact operands, so the nearest-mode bytes must be identical; bilinear is identical except where the
    interpolated value is within rounding of 0.5 (the golden carries each pixel's distance from the threshold)."""
    g = golden("yl_mask_binary")
    d = yl.dev
    for tag in "ab":
        proto, coeff, box = t(g[f"proto_{tag}"]).to(d), t(g[f"coeff_{tag}"]).to(d), t(g[f"box_{tag}"]).to(d)
        if simt:
            coeff = _off16(coeff)
        for j in range(5):
            size = tuple(int(v) for v in g[f"size_{tag}{j}"])
            near = yl.masks.assemble_mask_binary(proto, coeff, box, size)
            assert near.dtype == torch.uint8 and tuple(near.shape) == (coeff.shape[0],) + size
            if tag == "b" or simt:
                assert_equal(near, g[f"nearest_{tag}{j}"], f"nearest {tag}{j}")
            else:  # real-valued logits summed in another order: a pixel within rounding of logit 0 may flip
                lg = O.mask_logits(t(g[f"proto_{tag}"]), t(g[f"coeff_{tag}"]))
                iy, ix = O.upsample_nearest_index(size[0], lg.shape[1]), O.upsample_nearest_index(size[1], lg.shape[2])
                clear = (lg.abs() > 1e-4)[:, iy][:, :, ix].numpy()
                assert_equal(near.cpu().numpy()[clear], g[f"nearest_{tag}{j}"][clear], f"nearest {tag}{j}")
            bil = yl.masks.assemble_mask_binary(proto, coeff, box, size, mode="bilinear").cpu().numpy()
            clear = g[f"bilinear_dist_{tag}{j}"].astype(np.float32) > 2e-5
            assert clear.mean() > 0.99
            assert_equal(bil[clear], g[f"bilinear_{tag}{j}"][clear], f"bilinear {tag}{j}")
    e = yl.masks.assemble_mask_binary(proto, coeff[:0], box[:0], (11, 13))
    assert e.shape == (0, 11, 13) and e.dtype == torch.uint8
    with pytest.raises(ValueError):
        yl.masks.assemble_mask_binary(proto, coeff, box, (11, 13), mode="bicubic")


@pytest.mark.parametrize("H,W,K,ho,wo", [(138, 138, 100, 480, 640), (138, 138, 37, 550, 550), (69, 69, 150, 69, 69),
                                          (276, 276, 20, 241, 323), (69, 69, 60, 150, 212)])
@pytest.mark.parametrize("mode", ["nearest", "bilinear"])
def test_mask_binary_vs_oracle(yl, mode, H, W, K, ho, wo):
    """Full-size prototype maps, camera (480x640) and network-input (550x550) resolutions, identity and odd sizes
    (241x323 takes the byte-store path, 150x212 the 4-byte one, rows that are multiples of 16 the 128-bit one).  Every pixel whose resized fp32 mask value is clear of 0.5 must match the
    oracle; the rest (within the tensor-core contraction error of the threshold) are counted and must be rare."""
    proto, coeff, box = synth.mask_inputs(32, H, W, K, seed=H + K)
    d = yl.dev
    got = yl.masks.assemble_mask_binary(proto.to(d), coeff.to(d), box.to(d), (ho, wo), mode=mode).cpu()
    val = O.mask_upsampled(proto, coeff, box, (ho, wo), mode)
    clear = (val - 0.5).abs() > 5e-5
    assert float(clear.float().mean()) > 0.999
    assert_equal(got[clear], (val > 0.5).to(torch.uint8)[clear], f"{mode} {H}x{W} -> {ho}x{wo}")
    assert int(got.max()) == 1 and 0.01 < float(got.float().mean()) < 0.5  # (not vacuous)


def test_mask_binary_batched_matches_single(yl):
    d = yl.dev
    B, N, P, H, W, top_k, ho, wo = 3, 500, 32, 48, 52, 40, 120, 160
    g = synth.gen(12)
    anchor = torch.cat((torch.rand((1, N, 2), generator=g), torch.rand((1, N, 2), generator=g) * 0.3 + 0.05), -1)
    cls, enc = synth.yolact_heads(B, N, 7, seed=13, anchor=anchor, separated=True)
    proto = torch.nn.functional.leaky_relu(torch.randn((B, P, H, W), generator=g)).to(d)
    coeff = torch.tanh(torch.randn((B, N, P), generator=g)).to(d)
    det = yl.nms.detect(cls.to(d), enc.to(d), anchor.to(d), CFG, top_k, 0.5, 0.05)
    for mode in ("nearest", "bilinear"):
        out = torch.full((B, top_k, ho, wo), 7, dtype=torch.uint8, device=d)
        yl.masks.assemble_mask_binary_batched(proto, coeff, det, (ho, wo), mode=mode, out=out)
        for b in range(B):
            n = int(det.n_keep[b])
            assert n > 0
            single = yl.masks.assemble_mask_binary(proto[b], coeff[b, det.keep[b, :n]], det.box[b, :n], (ho, wo), mode=mode)
            assert_equal(out[b, :n], single, f"{mode} frame {b}: batched == per-frame call")
            assert bool((out[b, n:] == 7).all()), "rows beyond n_keep are left untouched"


def test_match_golden(yl):
    g = golden("yl_match")
    d = yl.dev
    m = yl.loss.match_anchors(t(g["anchor"]).to(d), t(g["truth_box"]).to(d), t(g["truth_valid"]).to(d), CFG)
    assert_equal(m.match_index, g["match_index"], "first max on ties"), assert_equal(m.match_iou, g["match_iou"])
    assert_equal(m.positive_match, g["positive"]), assert_equal(m.negative_match, g["negative"])
    assert_close(m.box_target[m.positive_match], g["targets"], atol=1e-6, what="regression targets of the positives")


@pytest.mark.parametrize("N,M", [(1000, 16), (77, 3), (4000, 40), (19248, 1)])
def test_match_culling_edge_cases_vs_oracle(yl, N, M):
    """The kernel culls, per warp of 32 consecutive priors, the truths that miss the warp's hull.  Truths / priors with
    NaN, infinite, zero-area or negative-size boxes must never be culled wrongly: index and IoU exact against the
    oracle (torch.max: first maximum wins, a NaN wins over what came before), N not a multiple of 32, M > 32."""
    g = synth.gen(100 + N + M)
    anchor = torch.cat((torch.rand((1, N, 2), generator=g), torch.rand((1, N, 2), generator=g) * 0.2 + 0.01), -1)
    tb, tv = synth.truth_boxes(3, M, seed=N)
    if M >= 3:
        tb[0, 1] = torch.tensor([float("nan"), 0.5, 0.1, 0.1])
        tb[1, 2] = torch.tensor([0.5, 0.5, 0.0, 0.0])              # zero area: 0/0 -> NaN only against ...
        tb[2, 0] = torch.tensor([0.5, float("inf"), 0.2, 0.2])
        tv[:, :3] = True
    if N >= 1000:
        anchor[0, 5] = torch.tensor([0.5, 0.5, 0.0, 0.0])          # ... a zero-area prior at the same spot
        anchor[0, 40] = torch.tensor([float("nan"), 0.2, 0.1, 0.1])
        anchor[0, 70, 2:] = torch.tensor([-0.1, 0.3])              # negative height
    mi, miou, pos, neg, _ = O.match_anchors(anchor, tb, tv, 0.4, 0.3, CFG.box_variances)
    d = yl.dev
    m = yl.loss.match_anchors(anchor.to(d), tb.to(d), tv.to(d), CFG)
    assert_equal(torch.isnan(m.match_iou.cpu()), torch.isnan(miou), "NaN positions of match_iou")
    assert_equal(torch.nan_to_num(m.match_iou.cpu(), nan=-7.0), torch.nan_to_num(miou, nan=-7.0), "match_iou")
    ok = ~torch.isnan(miou)   # (where the maximum is NaN its index is whichever NaN came first: compare those too)
    assert_equal(m.match_index.cpu()[ok], mi[ok]), assert_equal(m.match_index.cpu()[~ok], mi[~ok])
    assert_equal(m.positive_match, pos), assert_equal(m.negative_match, neg)


def test_match_full_size_vs_oracle(yl):
    anchor = O.all_anchors(synth.fpn_sizes(550, 550), CFG.anchor_scales, CFG.anchor_aspect_ratios, 550, 550)
    tb, tv = synth.truth_boxes(4, 16, seed=8)
    g = synth.gen(9)
    pick = torch.randint(0, 19248, (4, 8), generator=g)
    tb[:, :8] = anchor[0, pick] * (1 + 0.05 * torch.randn((4, 8, 4), generator=g).clamp(-1, 1) * torch.tensor([0, 0, 1, 1]))
    tv[:, :8] = True
    mi, miou, pos, neg, tgt = O.match_anchors(anchor, tb, tv, 0.4, 0.3, CFG.box_variances)
    d = yl.dev
    m = yl.loss.match_anchors(anchor.to(d), tb.to(d), tv.to(d), CFG)
    assert_equal(m.match_index, mi), assert_equal(m.match_iou, miou)
    assert_equal(m.positive_match, pos), assert_equal(m.negative_match, neg)
    assert int(pos.sum()) > 0
    assert_close(m.box_target[m.positive_match], tgt[pos], atol=1e-6, what="targets")


# ---- loss (YOLACT half) ----------------------------------------------------------------------------

def _loss_case(g, d):
    pred = tuple(t(g[k]).to(d) for k in ("cls", "enc", "coeff", "anchor", "proto"))
    truth = tuple(t(g[k]).to(d) for k in ("truth_valid", "truth_cls", "truth_box", "seg", "img_valid"))
    return pred, truth


def test_loss_golden(yl):
    """The reference's own loss values and autograd gradients (yolact/model/loss.py:8-125, tests/golden/make_golden.py):
    hard-negative mining, a frame without truths, a positive whose resized truth mask is empty."""
    g = golden("yl_loss")
    pred, truth = _loss_case(g, yl.dev)
    cls, enc, coeff, anchor, proto = pred
    for x in (cls, enc, coeff, proto):
        x.requires_grad_()
    total, (lc, lb, lm) = yl.loss.loss(pred, truth, CFG)
    assert_close(lc, g["cls_loss"], what="classification term"), assert_close(lb, g["box_loss"], what="box term")
    assert_close(lm, g["mask_loss"], what="mask term"), assert_close(total, g["total"], what="total")
    total.backward()
    assert_close(cls.grad, g["grad_cls"], atol=1e-9, what="d/d classification")
    assert_close(enc.grad, g["grad_enc"], atol=1e-9, what="d/d box_encoding")
    assert_close(coeff.grad, g["grad_coeff"], rtol=1e-4, atol=1e-8, what="d/d mask_coeff")
    assert_close(proto.grad, g["grad_proto"], rtol=1e-4, atol=1e-8, what="d/d mask_prototype")


@pytest.mark.parametrize("B,N,C1,M,ratio", [(2, 300, 81, 5, 3), (1, 19248, 81, 16, 3), (3, 1000, 4, 3, 0),
                                            (2, 2500, 33, 40, 7)])
def test_class_box_loss_vs_oracle(yl, B, N, C1, M, ratio):
    """Selected sets exact (random logits: no ties in the background confidence), losses and gradients 1e-5 against
    the oracle's autograd; N not a multiple of 32 or 1024, ratio 0 (no mining), many truths."""
    d = yl.dev
    g = synth.gen(7 * N + C1)
    anchor = torch.cat((torch.rand((1, N, 2), generator=g), torch.rand((1, N, 2), generator=g) * 0.2 + 0.02), -1)
    tb, tv = synth.truth_boxes(B, M, seed=N + 1)
    pick = torch.randint(0, N, (B, min(M, 3)), generator=g)
    tb[:, :pick.shape[1]] = anchor[0, pick] * (1 + 0.05 * torch.randn((B, pick.shape[1], 4), generator=g).clamp(-1, 1))
    tv[:, :pick.shape[1]] = True
    tcls = torch.randint(1, C1, (B, M), generator=g)
    cls = (torch.randn((B, N, C1), generator=g) * 3).requires_grad_()
    enc = (torch.randn((B, N, 4), generator=g) * 1.5).requires_grad_()
    cfg = SimpleNamespace(**{**vars(CFG), "negative_example_ratio": ratio})
    ocl, obl, osel = O.yolact_class_box_loss(cls, enc, anchor, tv, tcls, tb, cfg.iou_pos_threshold, cfg.iou_neg_threshold,
                                             cfg.box_variances, ratio)
    og_cls, og_enc = torch.autograd.grad(ocl + 2 * obl, (cls, enc))
    dcls, denc = cls.detach().to(d).requires_grad_(), enc.detach().to(d).requires_grad_()
    m = yl.loss.match_anchors(anchor.to(d), tb.to(d), tv.to(d), cfg)
    r = yl.loss.class_box_loss(dcls, denc, m, tcls.to(d), cfg)
    assert int(r.n_pos.sum()) > 0
    assert_equal(r.selected, osel, "positives + mined negatives")
    for b in range(B):
        n = int(r.n_pos[b])
        assert_equal(r.pos_list[b, :n], m.positive_match[b].nonzero().flatten().int(), "positives in prior order")
    assert_close(r.classification_loss, ocl.detach(), what="classification term")
    assert_close(r.box_loss, obl.detach(), what="box term")
    (r.classification_loss + 2 * r.box_loss).backward()
    assert_close(dcls.grad, og_cls, atol=1e-9, what="d/d classification")
    assert_close(denc.grad, og_enc, atol=1e-9, what="d/d box_encoding")


def test_class_box_loss_ties_and_short_negative_lists(yl):
    """Equal background confidences at the mining boundary go to the lower prior index, and when ratio * n_positive
    exceeds the frame's negatives the remaining picks are the lowest-index priors among the rest (the oracle's stable
    sort); quantised logits make the ties."""
    d = yl.dev
    g = synth.gen(55)
    B, N, C1, M = 2, 640, 5, 4
    anchor = torch.cat((torch.rand((1, N, 2), generator=g), torch.rand((1, N, 2), generator=g) * 0.2 + 0.02), -1)
    tb, tv = synth.truth_boxes(B, M, seed=56)
    tb[:, :2] = anchor[0, torch.randint(0, N, (B, 2), generator=g)]
    tv[:, :2] = True
    tcls = torch.randint(1, C1, (B, M), generator=g)
    cls = torch.round(torch.randn((B, N, C1), generator=g))          # integer logits: many equal softmax rows
    enc = torch.randn((B, N, 4), generator=g)
    for ratio in (3, 400):
        cfg = SimpleNamespace(**{**vars(CFG), "negative_example_ratio": ratio})
        ocl, obl, osel = O.yolact_class_box_loss(cls, enc, anchor, tv, tcls, tb, cfg.iou_pos_threshold,
                                                 cfg.iou_neg_threshold, cfg.box_variances, ratio)
        m = yl.loss.match_anchors(anchor.to(d), tb.to(d), tv.to(d), cfg)
        r = yl.loss.class_box_loss(cls.to(d), enc.to(d), m, tcls.to(d), cfg)
        assert_equal(r.selected, osel, f"ratio {ratio}")
        assert_close(r.classification_loss, ocl, what="classification term")


@pytest.mark.parametrize("B,N,K,M,PH,PW,SH,SW", [(2, 400, 32, 5, 138, 138, 550, 550), (1, 200, 12, 3, 17, 23, 40, 31),
                                                 (2, 300, 32, 4, 64, 64, 64, 64)])
def test_mask_loss_vs_oracle(yl, B, N, K, M, PH, PW, SH, SW):
    """Mask term and its gradients against the oracle's autograd: the YOLACT shapes (550 -> 138: non-integer resize
    scale), odd sizes with upsampling, identity resize; blobs of saturated logits exercise the clamps."""
    d = yl.dev
    g = synth.gen(31 * PH + K)
    anchor = torch.cat((torch.rand((1, N, 2), generator=g) * 0.6 + 0.2, torch.rand((1, N, 2), generator=g) * 0.3 + 0.1), -1)
    tb, tv = synth.truth_boxes(B, M, seed=PH)
    pick = torch.randint(0, N, (B, 2), generator=g)
    tb[:, :2] = anchor[0, pick] * (1 + 0.03 * torch.randn((B, 2, 4), generator=g).clamp(-1, 1))
    tv[:, :2] = True
    seg = torch.full((B, SH, SW), -1, dtype=torch.int64)
    for b in range(B):
        for j in range(M):
            y, x, h, w = (float(v) for v in tb[b, j])
            y0, y1 = max(int((y - h / 2) * SH), 0), min(int((y + h / 2) * SH) + 1, SH)
            x0, x1 = max(int((x - w / 2) * SW), 0), min(int((x + w / 2) * SW) + 1, SW)
            blob = torch.rand((max(y1 - y0, 0), max(x1 - x0, 0)), generator=g) < 0.7
            seg[b, y0:y1, x0:x1][blob] = j
    img_valid = torch.rand((B, SH, SW), generator=g) < 0.9
    coeff = (torch.randn((B, N, K), generator=g) * 1.5).requires_grad_()
    proto = (torch.randn((B, K, PH, PW), generator=g)).requires_grad_()
    cfg = CFG
    oml = O.yolact_mask_loss(coeff, proto, anchor, tv, tb, seg, img_valid, cfg.iou_pos_threshold, cfg.iou_neg_threshold,
                             cfg.box_variances)
    og_c, og_p = torch.autograd.grad(oml, (coeff, proto))
    m = yl.loss.match_anchors(anchor.to(d), tb.to(d), tv.to(d), cfg)
    n_pos = m.positive_match.sum(dim=1)
    assert int(n_pos.sum()) > 0
    pos_list = torch.zeros((B, N), dtype=torch.int32, device=d)
    for b in range(B):
        idx = m.positive_match[b].nonzero().flatten().int()
        pos_list[b, :idx.numel()] = idx
    dc, dp = coeff.detach().to(d).requires_grad_(), proto.detach().to(d).requires_grad_()
    ml = yl.loss.mask_loss(dc, dp, m, pos_list, n_pos, tb.to(d), seg.to(d), img_valid.to(d))
    assert_close(ml, oml.detach(), what="mask term")
    (3 * ml).backward()
    scale = float(og_c.abs().max())
    assert_close(dc.grad, 3 * og_c, rtol=1e-4, atol=1e-6 * scale, what="d/d mask_coeff")
    assert_close(dp.grad, 3 * og_p, rtol=1e-4, atol=1e-6 * float(og_p.abs().max()), what="d/d mask_prototype")
    again = yl.loss.mask_loss(dc, dp, m, pos_list, n_pos, tb.to(d), seg.to(d), img_valid.to(d))
    assert_equal(again, ml, "run-to-run")
    # the segmentation map is read in place in the caller's type: uint8 as the reference's dataset holds it
    # (segmentation_dataset.py:98-99; 255 = no object), int32, int64 — the same bits out of all three
    for dtype, none in ((torch.uint8, 255), (torch.int32, -1)):
        s = torch.where(seg < 0, torch.full_like(seg, none), seg).to(dtype).to(d)
        dc2, dp2 = coeff.detach().to(d).requires_grad_(), proto.detach().to(d).requires_grad_()
        ml2 = yl.loss.mask_loss(dc2, dp2, m, pos_list, n_pos, tb.to(d), s, img_valid.to(d))
        assert_equal(ml2, ml, f"seg as {dtype}")
        (3 * ml2).backward()
        assert_equal(dc2.grad, dc.grad, f"d/d mask_coeff, seg as {dtype}")
        assert_equal(dp2.grad, dp.grad, f"d/d mask_prototype, seg as {dtype}")


# ---- head outputs in the consumers' layout ----------------------------------------------------------

def test_pack_heads_golden(yl):
    """Against the reference's own PredictionHead (forward hooks on its final convolutions, tests/golden/make_golden.py):
    packed outputs bit-exact for the class logits and box encodings (pure data movement), tanh 1e-6; gradients back at
    the convolution outputs against the reference's autograd."""
    from tauv_vision_b200.yolact.model import prediction_head as PH
    g = golden("yl_heads")
    d, L = yl.dev, int(g["n_levels"])
    cfg = SimpleNamespace(n_classes=int(g["n_classes"]), n_prototype_masks=int(g["n_prototype_masks"]))
    lv = {k: [t(g[f"{k}_level{l}"]).to(d).requires_grad_() for l in range(L)] for k in ("cls", "box", "coeff")}
    cls, box, coeff = PH.pack_heads(lv["cls"], lv["box"], lv["coeff"], cfg)
    assert_equal(cls, g["cls"], "classification"), assert_equal(box, g["box"], "box_encoding")
    assert_close(coeff, g["coeff"], rtol=1e-6, atol=1e-7, what="mask_coeff (tanh)")
    ((cls * t(g["w_cls"]).to(d)).sum() + (box * t(g["w_box"]).to(d)).sum() + (coeff * t(g["w_coeff"]).to(d)).sum()).backward()
    for k in ("cls", "box"):
        for l in range(L):
            assert_equal(lv[k][l].grad, g[f"{k}_grad{l}"], f"d/d {k} level {l}")
    for l in range(L):
        assert_close(lv["coeff"][l].grad, g[f"coeff_grad{l}"], rtol=1e-5, atol=1e-7, what=f"d/d coeff level {l}")


def test_pack_heads_full_size_vs_oracle(yl):
    """The YOLACT shapes: five levels of 550 x 550 (69, 35, 18, 9, 5), three aspect ratios, 81 classes / 32
    coefficients; the packed rows line up with all_anchors' 19 248 priors."""
    from tauv_vision_b200.yolact.model import prediction_head as PH
    d = yl.dev
    g = synth.gen(77)
    sizes = synth.fpn_sizes(550, 550)
    for C, th in ((81, False), (4, False), (32, True)):
        lv = [torch.randn((2, 3 * C, h, w), generator=g) for h, w in sizes]
        want = O.pack_head(lv, C, tanh=th)
        got = PH.pack_head([x.to(d) for x in lv], C, tanh=th)
        assert got.shape == (2, 19248, C)
        if th:
            assert_close(got, want, rtol=1e-6, atol=1e-7, what="tanh")
        else:
            assert_equal(got, want, f"C = {C}")


def test_loss_is_run_to_run_identical(yl):
    """No floating-point atomics anywhere in the loss: two runs of loss() + backward on the same inputs give bit-identical
    values and gradients (fixed summation orders, fixed-point area sums, ordered compactions)."""
    d = yl.dev
    g = synth.gen(123)
    B, N, C1, K, M, PH, PW, SH, SW = 4, 3000, 21, 32, 6, 69, 69, 200, 200
    anchor = torch.cat((torch.rand((1, N, 2), generator=g) * 0.8 + 0.1, torch.rand((1, N, 2), generator=g) * 0.3 + 0.05), -1).to(d)
    tb, tv = synth.truth_boxes(B, M, seed=124)
    tb[:, :3] = anchor[0, torch.randint(0, N, (B, 3), generator=g)].cpu()
    tv[:, :3] = True
    truth = (tv.to(d), torch.randint(1, C1, (B, M), generator=g).to(d), tb.to(d),
             torch.randint(0, M, (B, SH, SW), generator=g).to(d), (torch.rand((B, SH, SW), generator=g) < 0.9).to(d))
    base = [torch.randn((B, N, C1), generator=g), torch.randn((B, N, 4), generator=g), torch.randn((B, N, K), generator=g),
            torch.randn((B, K, PH, PW), generator=g)]
    runs = []
    for _ in range(2):
        xs = [x.to(d).requires_grad_() for x in base]
        total, parts = yl.loss.loss((xs[0], xs[1], xs[2], anchor, xs[3]), truth, CFG)
        total.backward()
        runs.append([total.detach(), *[p.detach() for p in parts], *[x.grad for x in xs]])
    assert float(runs[0][3]) > 0 and float(runs[0][1]) > 0
    for a_, b_ in zip(*runs):
        assert torch.equal(a_, b_)
