"""Latency of the drop-in calls at batch 1 (the ROS-node case) and batch 64: device part vs Python list building."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from types import SimpleNamespace
from tauv_vision_b200.centernet.model import decode as D
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
for B in (1, 64):
    C, H, K = 80, 128, 100
    logits = torch.randn((B, C, H, H), device=dev, generator=g) * 1.5 - 2.2
    size = (torch.rand((B, 2, H, H), device=dev, generator=g) * 0.3).permute(0, 2, 3, 1)
    offset = (torch.rand((B, 2, H, H), device=dev, generator=g) * 4).permute(0, 2, 3, 1)
    pred = SimpleNamespace(heatmap=logits, size=size, offset=offset, depth=None)
    mc = SimpleNamespace(in_h=512, in_w=512, downsample_ratio=4, out_h=H, out_w=H)
    for thr in (0.3, 0.9):
        for _ in range(3): D.decode(pred, mc, K, thr)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20): out = D.decode(pred, mc, K, thr)
        t1 = time.perf_counter()
        for _ in range(20): p = D.decode_packed(pred, mc, K, thr); torch.cuda.synchronize()
        t2 = time.perf_counter()
        for _ in range(20): h = D.decode_packed(pred, mc, K, thr).to_host()
        t3 = time.perf_counter()
        n = sum(len(f) for f in out)
        print(f"B={B} thr={thr}: decode() {1e6*(t1-t0)/20:8.1f} us ({n} detections) | decode_packed+sync {1e6*(t2-t1)/20:8.1f} us | +to_host {1e6*(t3-t2)/20:8.1f} us")
