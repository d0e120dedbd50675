"""ncu / timing target: the fused CenterNet decode (one launch) on the bench workload, decode only, back to back.

    python tools/decode_once.py [reps] [B] [C] [H] [W] [K]
Prints the mean device time per launch over `reps` back-to-back launches (the previous launch only reads, so the L2
holds no dirty lines: the kernel's own streaming rate) and after a 512 MB memset (L2 full of dirty lines, as after
the target encode in bench.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import torch
import tauv_vision_b200 as tv
from tauv_vision_b200.centernet.model import decode as D
tv.load_library()
dev = torch.device("cuda", 0)
a = [int(x) for x in sys.argv[1:]]
reps = a[0] if len(a) > 0 else 20
B, C, H, W, K = (a[1:6] + [64, 80, 128, 128, 100][len(a[1:6]):]) if len(a) > 1 else (64, 80, 128, 128, 100)
g = torch.Generator(device=dev); g.manual_seed(1)
logits = torch.randn((B, C, H, W), device=dev, generator=g) * 1.5 - 2.2
size = (torch.rand((B, 2, H, W), device=dev, generator=g) * 0.3).permute(0, 2, 3, 1)
offset = (torch.rand((B, 2, H, W), device=dev, generator=g) * 4).permute(0, 2, 3, 1)
mc = SimpleNamespace(in_h=H * 4, in_w=W * 4, downsample_ratio=4, out_h=H, out_w=W)
pred = SimpleNamespace(heatmap=logits, size=size, offset=offset, depth=None)
out = D.decode_packed(pred, mc, K, 0.3)
for _ in range(3):
    D.decode_packed(pred, mc, K, 0.3, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    D.decode_packed(pred, mc, K, 0.3, out=out)
e1.record(); torch.cuda.synchronize()
t_clean = e0.elapsed_time(e1) / reps * 1e3
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
ts = []
for r in range(reps):
    flush.fill_(r & 255)
    e0.record(); D.decode_packed(pred, mc, K, 0.3, out=out); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
ts.sort()
nbytes = 4 * B * C * H * W
print(f"decode B={B} C={C} {H}x{W} k={K}: back-to-back {t_clean:.1f} us = {nbytes / t_clean / 1e3:.0f} GB/s; "
      f"after a 512 MB fill (dirty L2) median {ts[len(ts) // 2]:.1f} us = {nbytes / ts[len(ts) // 2] / 1e3:.0f} GB/s; "
      f"mean detections {float(out.count.float().mean()):.1f}")
