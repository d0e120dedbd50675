"""Timing target: the keypoint-affinity term fused with its target render at Kp = 80, 16 objects x 4 keypoints per frame,
64 frames of 128 x 128, against rendering the targets and composing the term with eager torch ops on the same GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import torch
from tauv_vision_b200.centernet.model import loss as L
from tests import synth
dev = torch.device("cuda", 0)
B, Kp, H, W, m = 64, 80, 128, 128, 64
tr = synth.truth_to(synth.pose_truth(B, 16, 20, seed=5, n_kp_inst=m, Kp=Kp), dev)
mc = SimpleNamespace(in_h=512, in_w=512, downsample_ratio=4, out_h=H, out_w=W)
tc = SimpleNamespace(keypoint_heatmap_sigma=2.0, keypoint_affinity_sigma=3.0)
pred = (torch.randn((B, Kp, 2, H, W), device=dev) * 0.5).requires_grad_()

def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); torch.cuda._sleep(2_000_000)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

with torch.no_grad():
    fwd = timed(lambda: L.keypoint_affinity_loss(pred, tr, mc, tc))
def fb():
    pred.grad = None
    L.keypoint_affinity_loss(pred, tr, mc, tc).backward()
both = timed(fb)
def eager():
    pred.grad = None
    _, w, t = L.generate_keypoint_heatmap(tr, mc, tc, SimpleNamespace(n_keypoints=Kp))
    (w.unsqueeze(2) * torch.nn.functional.mse_loss(pred, t, reduction="none")).sum().backward()
ref = timed(eager, n=3)
nb = 4 * B * Kp * 2 * H * W
print(f"keypoint affinity term {B}x{Kp}x2x{H}x{W}, {m} instances per frame: fused forward {fwd:.1f} us ({nb / fwd / 1e3:.0f} GB/s of "
      f"the predicted field), forward + backward {both:.1f} us; target render + eager torch ops (mse, weight, sum, autograd) "
      f"on this GPU {ref:.1f} us")
