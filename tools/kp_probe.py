"""Time generate_keypoint_heatmap at the configs[1] shape (B=64, Kp=80?, 128x128, 16 instances per frame)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from types import SimpleNamespace
from tauv_vision_b200.centernet.model import loss as L
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, n=7):
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
g = torch.Generator(device=dev); g.manual_seed(1)
for (B, Kp, H, n, m) in [(64, 80, 128, 16, 16), (64, 32, 128, 16, 32), (64, 24, 128, 16, 48)]:
    truth = SimpleNamespace(valid=torch.rand((B, n), device=dev, generator=g) < 0.75,
                            label=torch.randint(0, 8, (B, n), device=dev, generator=g),
                            center=torch.rand((B, n, 2), device=dev, generator=g),
                            keypoint_valid=torch.rand((B, m), device=dev, generator=g) < 0.75,
                            keypoint_label=torch.randint(0, Kp, (B, m), device=dev, generator=g),
                            keypoint_center=torch.rand((B, m, 2), device=dev, generator=g),
                            keypoint_object_index=torch.randint(0, n, (B, m), device=dev, generator=g))
    mc = SimpleNamespace(in_h=H * 4, in_w=H * 4, downsample_ratio=4, out_h=H, out_w=H)
    tc = SimpleNamespace(keypoint_heatmap_sigma=2.0, keypoint_affinity_sigma=4.0)
    oc = SimpleNamespace(n_labels=8, n_keypoints=Kp)
    t = timeit(lambda: L.generate_keypoint_heatmap(truth, mc, tc, oc))
    by = 4 * B * Kp * H * H * 4
    print(f"keypoint encode B={B} Kp={Kp} {H}x{H} m={m}: {t:.1f} us  {by / t / 1e3:.0f} GB/s written ({by/1e6:.0f} MB) frac {by / t / 1e3 / 6454.3:.2f}")
