"""Practical write roofline: torch fill / zero_ / memset of 335 MB, vs gaussian_encode."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from types import SimpleNamespace
from tauv_vision_b200.centernet.model import loss as L
dev = torch.device("cuda", 0)
B, C, H, W = 64, 80, 128, 128
out = torch.empty((B, C, H, W), device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, n=10, do_flush=True):
    ts = []
    for _ in range(n):
        if do_flush: flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
gb = out.numel() * 4 / 1e9
for name, fn in (("zero_", lambda: out.zero_()), ("fill_(1.5)", lambda: out.fill_(1.5))):
    t = timeit(fn); print(f"{name:14s} {t:7.1f} us  {gb / t * 1e6:6.0f} GB/s")
g = torch.Generator(device=dev); g.manual_seed(1)
truth = SimpleNamespace(valid=torch.rand((B, 16), device=dev, generator=g) < 0.75,
                        label=torch.randint(0, C, (B, 16), device=dev, generator=g),
                        center=torch.rand((B, 16, 2), device=dev, generator=g))
mc = SimpleNamespace(in_h=512, in_w=512, downsample_ratio=4, out_h=H, out_w=W)
tc = SimpleNamespace(keypoint_heatmap_sigma=2.0); oc = SimpleNamespace(n_labels=C)
t = timeit(lambda: L.generate_heatmap(truth, mc, tc, oc, out=out)); print(f"{'gaussian_encode':14s} {t:7.1f} us  {gb / t * 1e6:6.0f} GB/s")
t = timeit(lambda: L.generate_heatmap(truth, mc, tc, oc, out=out), do_flush=False); print(f"{'  (no flush)':14s} {t:7.1f} us  {gb / t * 1e6:6.0f} GB/s")
