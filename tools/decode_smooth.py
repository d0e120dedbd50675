"""Decode time on SMOOTH heat maps (box-filtered noise, a few percent of the cells are peaks and the hot regions span many
blocks — what a trained CenterNet head emits), where the first threshold of the select pass finds too few peaks:
    python tools/decode_smooth.py [box] [passes]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import torch
from tauv_vision_b200.centernet.model import decode as D
dev = torch.device("cuda", 0)
box = int(sys.argv[1]) if len(sys.argv) > 1 else 5
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 1
B, C, H, W, K = 64, 80, 128, 128, 100
g = torch.Generator(device=dev); g.manual_seed(1)
x = torch.randn((B, C, H + passes * (box - 1), W + passes * (box - 1)), device=dev, generator=g)
for _ in range(passes):
    x = torch.nn.functional.avg_pool2d(x, box, 1)
x = ((x - x.mean()) / x.std() * 1.5 - 2.2).contiguous()
size = (torch.rand((B, 2, H, W), device=dev, generator=g) * 0.3).permute(0, 2, 3, 1)
offset = (torch.rand((B, 2, H, W), device=dev, generator=g) * 4).permute(0, 2, 3, 1)
mc = SimpleNamespace(in_h=H * 4, in_w=W * 4, downsample_ratio=4, out_h=H, out_w=W)
pred = SimpleNamespace(heatmap=x, size=size, offset=offset, depth=None)
s = torch.sigmoid(x); peaks = (torch.nn.functional.max_pool2d(s, 3, 1, 1) == s).float().mean().item()
out = D.decode_packed(pred, mc, K, 0.3)
for _ in range(3): D.decode_packed(pred, mc, K, 0.3, out=out)
torch.cuda.synchronize(); torch.cuda._sleep(1_000_000)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): D.decode_packed(pred, mc, K, 0.3, out=out)
e1.record(); torch.cuda.synchronize()
print(f"smooth maps (box {box} x {passes}): {100 * peaks:.2f} % of the cells are 3x3 peaks; decode {e0.elapsed_time(e1) * 100:.1f} us per 64 frames")
