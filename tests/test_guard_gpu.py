"""GPU: out-of-bounds WRITE check of the round-2 entry points (compute-sanitizer is not available on the GPU pool): every
output and workspace buffer is a window inside a larger allocation whose margins hold a sentinel; after the call through
the C ABI (odd sizes: N not a multiple of 32 / 1024, H*W not a multiple of the tile, K < 32) the margins must be intact."""
import ctypes
from types import SimpleNamespace

import pytest
import torch

from tests import synth

pytestmark = pytest.mark.gpu

GUARD = 4096  # bytes on either side


class Guarded:
    def __init__(self, shape, dtype, dev):
        n = 1
        for s in shape:
            n *= int(s)
        self.nbytes = n * torch.empty((), dtype=dtype).element_size()
        self.raw = torch.full((self.nbytes + 2 * GUARD + 256,), 0xA5, dtype=torch.uint8, device=dev)
        off = GUARD + (-(self.raw.data_ptr() + GUARD)) % 256          # 256-byte aligned window
        self.off = off
        self.t = self.raw[off:off + self.nbytes].view(dtype).reshape(shape)

    def check(self, what):
        lo, hi = self.raw[:self.off], self.raw[self.off + self.nbytes:]
        assert bool((lo == 0xA5).all()) and bool((hi == 0xA5).all()), f"{what}: write outside the buffer"


@pytest.fixture(scope="module")
def lib(cuda_device):
    import tauv_vision_b200 as tv
    return tv.load_library()


def _ptr(t, ctype):
    return ctypes.cast(t.data_ptr(), ctypes.POINTER(ctype))


def _f(t):
    return _ptr(t, ctypes.c_float)


def test_yolact_loss_entries_stay_inside_their_buffers(cuda_device, lib):
    from tauv_vision_b200 import _lib
    from tauv_vision_b200.yolact.model import loss as YL
    d = cuda_device
    B, N, C1, M, K, PH, PW, SH, SW = 3, 1237, 7, 5, 12, 19, 23, 41, 37
    g = synth.gen(5)
    cfg = synth.yolact_config()
    anchor = torch.cat((torch.rand((1, N, 2), generator=g), torch.rand((1, N, 2), generator=g) * 0.2 + 0.02), -1).to(d)
    tb, tv_ = synth.truth_boxes(B, M, seed=6)
    tb[:, :2] = anchor[0, torch.randint(0, N, (B, 2), generator=g)].cpu()
    tv_[:, :2] = True
    tb, tv_ = tb.to(d), tv_.to(d)
    m = YL.match_anchors(anchor, tb, tv_, cfg)
    assert int(m.positive_match.sum()) > 0
    cls = torch.randn((B, N, C1), generator=g).to(d)
    enc = torch.randn((B, N, 4), generator=g).to(d)
    tcls = torch.randint(1, C1, (B, M), generator=g).to(d)
    pos, neg = m.positive_match.view(torch.uint8), m.negative_match.view(torch.uint8)
    sel, pl = Guarded((B, N), torch.uint8, d), Guarded((B, N), torch.int32, d)
    sums, npos = Guarded((B, 2), torch.float64, d), Guarded((B,), torch.int64, d)
    ws = Guarded((lib.tauv_yolact_class_box_loss_workspace_bytes(B, N),), torch.uint8, d)
    _lib.check(lib.tauv_yolact_class_box_loss(
        _f(cls), _f(enc), _f(m.box_target), _ptr(pos, ctypes.c_uint8), _ptr(neg, ctypes.c_uint8),
        _ptr(m.match_index, ctypes.c_int64), _ptr(tcls, ctypes.c_int64), B, N, C1, M, 3, _ptr(sel.t, ctypes.c_uint8),
        _ptr(pl.t, ctypes.c_int32), _ptr(sums.t, ctypes.c_double), _ptr(npos.t, ctypes.c_int64), ws.t.data_ptr(),
        ws.t.numel(), _lib.stream_ptr(d)))
    torch.cuda.synchronize()
    for b_, w_ in ((sel, "selected"), (pl, "pos_list"), (sums, "sums"), (npos, "n_pos"), (ws, "workspace")):
        b_.check(w_)
    P = npos.t.sum().reshape(1)
    one = torch.ones(1, device=d)
    gcls, genc = Guarded((B, N, C1), torch.float32, d), Guarded((B, N, 4), torch.float32, d)
    _lib.check(lib.tauv_yolact_class_box_loss_backward(
        _f(cls), _f(enc), _f(m.box_target), _ptr(pos, ctypes.c_uint8), _ptr(sel.t, ctypes.c_uint8),
        _ptr(m.match_index, ctypes.c_int64), _ptr(tcls, ctypes.c_int64), B, N, C1, M, 3, _ptr(P, ctypes.c_int64), _f(one),
        _f(one), _f(gcls.t), _f(genc.t), _lib.stream_ptr(d)))
    torch.cuda.synchronize()
    gcls.check("grad_cls"), genc.check("grad_enc")
    # mask term
    coeff = torch.randn((B, N, K), generator=g).to(d)
    proto = torch.randn((B, K, PH, PW), generator=g).to(d)
    seg = torch.randint(-1, M, (B, SH, SW), generator=g).to(torch.int32).to(d)
    valid = (torch.rand((B, SH, SW), generator=g) < 0.9).to(torch.uint8).to(d)
    tsum = Guarded((B, M), torch.float64, d)
    part = Guarded((B, lib.tauv_yolact_mask_loss_partials()), torch.float64, d)
    recs = Guarded((lib.tauv_yolact_mask_loss_records_bytes(B, N),), torch.uint8, d)
    args = (_f(coeff), _f(proto), _ptr(pl.t, ctypes.c_int32), _ptr(npos.t, ctypes.c_int64), _ptr(m.match_index, ctypes.c_int64),
            _f(tb), seg.data_ptr(), 4, _ptr(valid, ctypes.c_uint8), B, N, K, M, PH, PW, SH, SW)
    _lib.check(lib.tauv_yolact_mask_loss(*args, _ptr(tsum.t, ctypes.c_double), recs.t.data_ptr(), _ptr(part.t, ctypes.c_double),
                                         _lib.stream_ptr(d)))
    torch.cuda.synchronize()
    tsum.check("tsum"), part.check("partial"), recs.check("records")
    gco, gpr = Guarded((B, N, K), torch.float32, d), Guarded((B, K, PH, PW), torch.float32, d)
    _lib.check(lib.tauv_yolact_mask_loss_backward(*args, _ptr(tsum.t, ctypes.c_double), recs.t.data_ptr(), _ptr(P, ctypes.c_int64), _f(one),
                                                  _f(gco.t), _f(gpr.t), _lib.stream_ptr(d)))
    torch.cuda.synchronize()
    gco.check("grad_coeff"), gpr.check("grad_proto")
    assert torch.isfinite(gco.t).all() and torch.isfinite(gpr.t).all()


def test_pack_heads_stays_inside_its_buffers(cuda_device, lib):
    from tauv_vision_b200 import _lib
    d = cuda_device
    B, CH = 2, 3 * 7
    sizes = [(9, 7), (5, 4), (33, 2), (1, 1)]
    g = synth.gen(9)
    levels = [torch.randn((B, CH, h, w), generator=g).to(d) for h, w in sizes]
    hw = (ctypes.c_int32 * len(sizes))(*[h * w for h, w in sizes])
    rows = sum(h * w for h, w in sizes)
    out = Guarded((B, rows * CH), torch.float32, d)
    ptrs = (ctypes.c_void_p * len(levels))(*[t_.data_ptr() for t_ in levels])
    _lib.check(lib.tauv_yolact_pack_heads(ptrs, hw, len(levels), B, CH, 1, _f(out.t), _lib.stream_ptr(d)))
    torch.cuda.synchronize()
    out.check("packed")
    grads = [Guarded((B, CH, h, w), torch.float32, d) for h, w in sizes]
    gptrs = (ctypes.c_void_p * len(levels))(*[g_.t.data_ptr() for g_ in grads])
    go = torch.randn((B, rows * CH), generator=g).to(d)
    _lib.check(lib.tauv_yolact_pack_heads_backward(_f(go), _f(out.t), hw, len(levels), B, CH, 1, gptrs, _lib.stream_ptr(d)))
    torch.cuda.synchronize()
    for l, g_ in enumerate(grads):
        g_.check(f"grad level {l}")


def test_centernet_loss_entries_stay_inside_their_buffers(cuda_device, lib):
    from tauv_vision_b200 import _lib
    d = cuda_device
    B, n_obj, Kp, m, H, W = 2, 5, 7, 41, 20, 36
    tr = synth.truth_to(synth.pose_truth(B, n_obj, 3, seed=3, n_kp_inst=m, Kp=Kp), d)
    pred = torch.randn((B, Kp, 2, H, W), generator=synth.gen(4)).to(d)
    kv = tr.keypoint_valid.view(torch.uint8)
    part = Guarded((lib.tauv_keypoint_affinity_loss_partials(B, Kp, H, W),), torch.float64, d)
    common = (_f(pred), _ptr(kv, ctypes.c_uint8), _ptr(tr.keypoint_label, ctypes.c_int64), _f(tr.keypoint_center),
              _ptr(tr.keypoint_object_index, ctypes.c_int64), _f(tr.center), B, m, n_obj, Kp, H, W, 4 * H, 4 * W, 4, 3.0)
    _lib.check(lib.tauv_keypoint_affinity_loss(*common, _ptr(part.t, ctypes.c_double), _lib.stream_ptr(d)))
    torch.cuda.synchronize()
    part.check("affinity partials")
    grad = Guarded((B, Kp, 2, H, W), torch.float32, d)
    one = torch.ones(1, device=d)
    _lib.check(lib.tauv_keypoint_affinity_loss_backward(*common, _f(one), _f(grad.t), _lib.stream_ptr(d)))
    torch.cuda.synchronize()
    grad.check("affinity gradient")
    # gather / scatter at the objects' cells
    idx = torch.stack((torch.randint(0, H, (B, n_obj)), torch.randint(0, W, (B, n_obj))), -1).to(d)
    C = 3
    dst = Guarded((B, C, H, W), torch.float32, d)
    dst.t.zero_()
    go = torch.randn((B, n_obj, C), generator=synth.gen(8)).to(d)
    _lib.check(lib.tauv_scatter_add_at(_f(go), _ptr(idx, ctypes.c_int64), B, n_obj, C, _f(dst.t), C * H * W, H * W, W, 1,
                                       _lib.stream_ptr(d)))
    torch.cuda.synchronize()
    dst.check("scatter_add_at")


@pytest.mark.parametrize("smooth", [False, True])
def test_decode_outputs_stay_inside_their_buffers(cuda_device, smooth):
    """The two-launch decode (block maxima + select) on noise and on smooth maps (the second threshold, the in-block
    filter, the compacted candidate list): every output tensor is a guarded window."""
    from tauv_vision_b200.centernet.model import decode as D
    d = cuda_device
    B, C, H, W, k = 3, 8, 64, 64, 100
    g = torch.Generator(device=d)
    g.manual_seed(3)
    x = torch.randn((B, C, H + (16 if smooth else 0), W + (16 if smooth else 0)), device=d, generator=g)
    if smooth:
        for _ in range(2):
            x = torch.nn.functional.avg_pool2d(x, 9, 1)
    x = ((x - x.mean()) / x.std() * 1.5 - 2.2).contiguous()
    size = (torch.rand((B, 2, H, W), device=d, generator=g) * 0.3).permute(0, 2, 3, 1)
    offset = (torch.rand((B, 2, H, W), device=d, generator=g) * 4).permute(0, 2, 3, 1)
    mc = SimpleNamespace(in_h=H * 4, in_w=W * 4, downsample_ratio=4, out_h=H, out_w=W)
    pred = SimpleNamespace(heatmap=x, size=size, offset=offset, depth=None)
    want = D.decode_packed(pred, mc, k, 0.3)
    bufs = {"index": Guarded((B, k, 2), torch.int64, d), "label": Guarded((B, k), torch.int64, d),
            "score": Guarded((B, k), torch.float32, d), "yx": Guarded((B, k, 2), torch.float64, d),
            "hw": Guarded((B, k, 2), torch.float32, d), "count": Guarded((B,), torch.int32, d)}
    out = D.PackedDetections(index=bufs["index"].t, label=bufs["label"].t, score=bufs["score"].t, yx=bufs["yx"].t,
                             hw=bufs["hw"].t, depth=None, count=bufs["count"].t)
    got = D.decode_packed(pred, mc, k, 0.3, out=out)
    torch.cuda.synchronize()
    for name, b_ in bufs.items():
        b_.check(name)
        assert torch.equal(getattr(got, name), getattr(want, name)), name
