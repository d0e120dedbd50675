"""Timing target: the YOLACT loss (match + class/box terms + mask term, forward and forward + backward) at BASELINE
configs[2]'s shapes (B = 64, 19 248 priors, 81 classes, 32 prototypes of 138 x 138, 550 x 550 segmentation maps,
16 truths per frame), against the reference's own op sequence (oracle/ref_port.py's restatement of yolact/model/loss.py,
a Python loop over frames and positives) run as eager torch ops on the same GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tauv_vision_b200.yolact.model import anchors as A, loss as L
from tests import synth
dev = torch.device("cuda", 0)
B, C1, K, M, PH, SH = (int(v) for v in (sys.argv[1:] + [64, 81, 32, 16, 138, 550][len(sys.argv) - 1:]))
cfg = synth.yolact_config()
anchor = A.all_anchors(synth.fpn_sizes(550, 550), cfg, dev)
N = anchor.shape[1]
g = torch.Generator(device=dev); g.manual_seed(3)
tb, tv = synth.truth_boxes(B, M, seed=4)
pick = torch.randint(0, N, (B, M // 2), generator=torch.Generator().manual_seed(5))
tb[:, :M // 2] = anchor[0].cpu()[pick] * (1 + 0.05 * torch.randn((B, M // 2, 4), generator=torch.Generator().manual_seed(6)).clamp(-1, 1))
tv[:, :M // 2] = True
tb, tv = tb.to(dev), tv.to(dev)
tcls = torch.randint(1, C1, (B, M), device=dev, generator=g)
seg = torch.randint(0, M, (B, SH // 10, SH // 10), device=dev, generator=g, dtype=torch.uint8).repeat_interleave(10, 1).repeat_interleave(10, 2)[:, :SH, :SH].contiguous()
valid = torch.ones((B, SH, SH), dtype=torch.bool, device=dev)
cls = (torch.randn((B, N, C1), device=dev, generator=g) * 2).requires_grad_()
enc = (torch.randn((B, N, 4), device=dev, generator=g)).requires_grad_()
coeff = torch.tanh(torch.randn((B, N, K), device=dev, generator=g)).requires_grad_()
proto = torch.relu(torch.randn((B, K, PH, PH), device=dev, generator=g)).requires_grad_()
pred, truth = (cls, enc, coeff, anchor, proto), (tv, tcls, tb, seg, valid)

def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); torch.cuda._sleep(2_000_000)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

def parts():
    m = L.match_anchors(anchor, tb, tv, cfg)
    cb = L.class_box_loss(cls, enc, m, tcls, cfg)
    return m, cb
m, cb = parts()
n_pos = int(cb.n_pos.sum())
t_match = timed(lambda: L.match_anchors(anchor, tb, tv, cfg))
with torch.no_grad():
    t_cb = timed(lambda: L.class_box_loss(cls, enc, m, tcls, cfg))
    t_mask = timed(lambda: L.mask_loss(coeff, proto, m, cb.pos_list, cb.n_pos, tb, seg, valid))
    t_fwd = timed(lambda: L.loss(pred, truth, cfg))
def fb():
    for x in (cls, enc, coeff, proto): x.grad = None
    L.loss(pred, truth, cfg)[0].backward()
t_both = timed(fb)
print(f"YOLACT loss, B={B} N={N} C1={C1} K={K} proto {PH}x{PH} seg {SH}x{SH}, {n_pos} positives ({n_pos / B:.1f} per frame): "
      f"match {t_match:.0f} us, class+box terms {t_cb:.0f} us ({4 * B * N * C1 / t_cb / 1e3:.0f} GB/s of class logits), "
      f"mask term {t_mask:.0f} us; loss() forward {t_fwd:.0f} us, forward + backward {t_both:.0f} us")
if os.environ.get("EAGER", "1") == "1":
    from oracle import ref_port as O   # (the checker's op sequence, timed as the eager-GPU baseline; not the product)
    def eager():
        for x in (cls, enc, coeff, proto): x.grad = None
        cl, bl, _ = O.yolact_class_box_loss(cls, enc, anchor, tv, tcls, tb, 0.4, 0.3, (0.1, 0.2), 3)
        ml = O.yolact_mask_loss(coeff, proto, anchor, tv, tb, seg, valid, 0.4, 0.3, (0.1, 0.2))
        (cl + bl + ml).backward()
    try:
        t_ref = timed(eager, n=1)
        print(f"the reference's op sequence as eager torch ops on this GPU, forward + backward: {t_ref / 1e3:.1f} ms")
    except Exception as e:  # noqa: BLE001
        print("eager baseline not run:", type(e).__name__, e)
