"""Timeline of the fused tail of tile_cluster_kernel (whole CenterNet decode in one launch), per CTA."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from types import SimpleNamespace
import tauv_vision_b200 as tv
from tauv_vision_b200.centernet.model import decode as D
lib = tv.load_library()
dev = torch.device("cuda", 0)
B, C, H, W, K = 64, 80, 128, 128, 100
g = torch.Generator(device=dev); g.manual_seed(1)
logits = torch.randn((B, C, H, W), device=dev, generator=g) * 1.5 - 2.2
size = torch.rand((B, 2, H, W), device=dev, generator=g).permute(0, 2, 3, 1)
offset = torch.rand((B, 2, H, W), device=dev, generator=g).permute(0, 2, 3, 1)
pred = SimpleNamespace(heatmap=logits, size=size, offset=offset, depth=None)
cfg = SimpleNamespace(in_h=512, in_w=512, downsample_ratio=4, out_h=H, out_w=W)
out = D.PackedDetections.allocate(B, K, False, dev)
run = lambda: D.decode_packed(pred, cfg, K, 0.3, out=out)
run(); run(); torch.cuda.synchronize()
n_items = B * C
trace = torch.zeros((n_items, 8), dtype=torch.int64, device=dev)
lib.tauv_debug_tile_trace.argtypes = [ctypes.c_void_p]
lib.tauv_debug_tile_trace(trace.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
lib.tauv_debug_tile_trace(None)
t = trace.cpu().numpy().astype(np.float64).reshape(B, C, 8)
t0 = t[:, 8:16, 0][t[:, 8:16, 0] > 0].min()          # first streaming item start anywhere
us = lambda x: (x - t0) / 1e3
second, third, last = t[:, 8:16, :], t[:, 16:24, :], t[:, -8:, :]
stream_end = us(second[:, :, 3])
s = [us(second[:, :, 1]), us(second[:, :, 6]), us(second[:, :, 7]), us(third[:, :, 1]), us(third[:, :, 6]), us(third[:, :, 7])]
names = ["sync 3a (all CTAs streamed + converted)", "pruned", "sync 3b", "pool copied + ranked", "outputs written", "sync 3 (unit done)"]
print(f"kernel (events) {e0.elapsed_time(e1)*1e3:.1f} us; kernel start at {us(last[:, :, 6]).min():.2f}")
print(f"stream end per CTA: min {stream_end.min():.1f} mean {stream_end.mean():.1f} max {stream_end.max():.1f}")
prev = stream_end
for n, x in zip(names, s):
    print(f"  {n:42s} +{(x - prev).mean():5.2f} us   (at mean {x.mean():6.1f}, max {x.max():6.1f})")
    prev = x
