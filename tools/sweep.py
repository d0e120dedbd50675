"""Throughput sweep over the BASELINE.json parity/sweep configs (#2 multi-scale, #5 classes x top-k, YOLACT priors x top_k).
Device-resident inputs, CUDA events, L2 flushed (by a read) between repetitions.  Prints a markdown table (profiles/r2_sweep.md)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from types import SimpleNamespace
import tauv_vision_b200 as tv
from tauv_vision_b200.centernet.model import decode as D, loss as L
from tauv_vision_b200.yolact.model import nms, anchors
from tests import synth

dev = torch.device("cuda", 0)
flush = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
PEAK = 6454.3

def timeit(fn, n=7):
    """Median device time of one call.  Between repetitions the L2 is flushed by READING 256 MB (a fill would leave it
    full of dirty lines, whose write-back the next kernel then pays for), and a short device-side sleep lets the host
    enqueue the call before the first event completes (otherwise small shapes time the host's launch path)."""
    ts = []
    for _ in range(n):
        flush.sum()
        torch.cuda._sleep(200000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]

rows = []
g = torch.Generator(device=dev); g.manual_seed(7)
B = 64
print("| workload | shape | time (us) | frames/s | algorithmic GB/s | frac of 6454 GB/s |\n|---|---|---|---|---|---|")
for (C, HW, K) in [(80, 256, 100), (80, 128, 100), (80, 64, 100), (80, 128, 200), (80, 128, 500), (80, 128, 1000),
                   (16, 128, 100), (4, 128, 100), (1, 128, 100), (1, 128, 1000)]:
    ratio = 512 // HW
    logits = torch.randn((B, C, HW, HW), device=dev, generator=g) * 1.5 - 2.2
    size = (torch.rand((B, 2, HW, HW), device=dev, generator=g) * 0.3).permute(0, 2, 3, 1)
    offset = (torch.rand((B, 2, HW, HW), device=dev, generator=g) * ratio).permute(0, 2, 3, 1)
    pred = SimpleNamespace(heatmap=logits, size=size, offset=offset, depth=None)
    mc = SimpleNamespace(in_h=512, in_w=512, downsample_ratio=ratio, out_h=HW, out_w=HW)
    buf = D.decode_packed(pred, mc, K, 0.3)
    t = timeit(lambda: D.decode_packed(pred, mc, K, 0.3, out=buf))
    by = 4 * B * C * HW * HW + B * K * 68
    print(f"| CenterNet decode | B={B} C={C} {HW}x{HW} k={K} | {t:.1f} | {B / t * 1e6:,.0f} | {by / t / 1e3:.0f} | {by / t / 1e3 / PEAK:.2f} |")
    if K == 100 and C == 80:
        truth = SimpleNamespace(valid=torch.rand((B, 16), device=dev, generator=g) < 0.75,
                                label=torch.randint(0, C, (B, 16), device=dev, generator=g),
                                center=torch.rand((B, 16, 2), device=dev, generator=g))
        out = torch.empty((B, C, HW, HW), device=dev)
        tc = SimpleNamespace(keypoint_heatmap_sigma=2.0); oc = SimpleNamespace(n_labels=C)
        t = timeit(lambda: L.generate_heatmap(truth, mc, tc, oc, out=out))
        by = 4 * B * C * HW * HW
        print(f"| Gaussian target encode | B={B} C={C} {HW}x{HW} n=16 | {t:.1f} | {B / t * 1e6:,.0f} | {by / t / 1e3:.0f} | {by / t / 1e3 / PEAK:.2f} |")
    del logits, size, offset, pred, buf
    torch.cuda.empty_cache()

cfg = synth.yolact_config()
for (N, TOPK) in [(19248, 100), (19248, 200), (19248, 500), (19248, 1000), (14505, 200), (4835, 200)]:
    C1 = 81
    cls = torch.randn((B, N, C1), device=dev, generator=g) * 2
    cls[:, :, 0] += 4
    idx = torch.randint(0, N - 16, (B, 12), device=dev, generator=g)
    for j in range(12):
        for o in range(12):
            cls[torch.arange(B, device=dev), idx[:, j] + o, 1 + (j % (C1 - 1))] += 10 + torch.rand((B,), device=dev, generator=g) * 4
    enc = torch.randn((B, N, 4), device=dev, generator=g) * 0.3
    anchor = torch.cat((torch.rand((1, N, 2), device=dev, generator=g), torch.rand((1, N, 2), device=dev, generator=g) * 0.3 + 0.05), -1)
    t = timeit(lambda: nms.detect(cls, enc, anchor, cfg, TOPK, 0.5, 0.05))
    by = B * N * (4 * C1 + 16) + 16 * N
    print(f"| YOLACT detect (scores + top_k + decode + Fast NMS) | B={B} N={N} C+1={C1} top_k={TOPK} | {t:.1f} | {B / t * 1e6:,.0f} | {by / t / 1e3:.0f} | {by / t / 1e3 / PEAK:.2f} |")
    del cls, enc
    torch.cuda.empty_cache()
