"""ncu / timing target: anchor matching at BASELINE configs[2] (64 frames, 19 248 priors, 16 truths per frame)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tauv_vision_b200.yolact.model import anchors, loss as yl_loss
from tests import synth
dev = torch.device("cuda", 0)
cfg = synth.yolact_config()
anchor = anchors.all_anchors(synth.fpn_sizes(550, 550), cfg, dev)
tb, tv = synth.truth_boxes(64, 16, seed=1)
tb, tv = tb.to(dev), tv.to(dev)
big = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for _ in range(9):
    big.sum()  # evict the outputs of the previous launch without leaving dirty lines
    torch.cuda._sleep(400000)  # (device-side spin: the wrapper's allocations and launch are enqueued meanwhile)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = yl_loss.match_anchors(anchor, tb, tv, cfg); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
ts.sort()
print(f"match_anchors 64 x 19248 x 16: median {ts[len(ts) // 2]:.1f} us (min {ts[0]:.1f})")
