"""Timing target: the fused heatmap focal loss (forward, and forward + backward) at BASELINE configs[1]."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import torch
from tauv_vision_b200.centernet.model import loss as L
from tests import synth
dev = torch.device("cuda", 0)
B, C, H, W = 64, 80, 128, 128
torch.manual_seed(0)
logits = (torch.randn((B, C, H, W), device=dev) * 1.5 - 2.2)
tr = synth.truth_to(synth.pose_truth(B, 16, C, seed=5), dev)
mc = SimpleNamespace(in_h=512, in_w=512, downsample_ratio=4, out_h=H, out_w=W)
tc = SimpleNamespace(keypoint_heatmap_sigma=2.0, heatmap_focal_loss_a=2.0, heatmap_focal_loss_b=4.0)

def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); torch.cuda._sleep(2_000_000)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

fwd = timed(lambda: L.heatmap_focal_loss(logits, tr, mc, tc))
x = logits.clone().requires_grad_(True)
def fb():
    x.grad = None
    L.heatmap_focal_loss(x, tr, mc, tc).backward()
both = timed(fb)
def eager():
    x.grad = None
    t = L.generate_heatmap(tr, mc, tc, SimpleNamespace(n_labels=C))
    L.focal_loss(torch.sigmoid(x), t, 2.0, 4.0).sum().backward()
ref = timed(eager, n=3)
nb = 4 * B * C * H * W
print(f"heatmap focal loss 64x80x128x128: fused forward {fwd:.1f} us ({nb / fwd / 1e3:.0f} GB/s of logits), forward+backward {both:.1f} us; "
      f"the same arithmetic as eager torch ops on this GPU (target render + ~20 elementwise passes + autograd) {ref:.1f} us")
