"""One batch of the fused mask + depth-mean path (for `ncu --metrics gpu__time_duration.sum` launch lists)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from types import SimpleNamespace
from tauv_vision_b200.yolact.model import masks
dev = torch.device("cuda", 0)
B, N, P, HP, TOPK, HI, WI = int(os.environ.get("B", 64)), 19248, 32, 276, 200, 720, 1280
g = torch.Generator(device=dev); g.manual_seed(3)
coeff = torch.tanh(torch.randn((B, N, P), device=dev, generator=g))
proto = torch.nn.functional.leaky_relu(torch.randn((B, P, HP, HP), device=dev, generator=g))
keep = torch.randint(0, N, (B, TOPK), device=dev, generator=g)
n_keep = torch.full((B,), 160, dtype=torch.int32, device=dev)
box = torch.cat((torch.rand((B, TOPK, 2), device=dev, generator=g) * 0.8 + 0.1, torch.rand((B, TOPK, 2), device=dev, generator=g) * 0.4 + 0.05), -1)
det = SimpleNamespace(keep=keep, n_keep=n_keep, box=box)
depth = torch.randint(300, 9000, (B, HI, WI), device=dev, dtype=torch.int32).to(torch.uint16)
for _ in range(3):
    masks.masked_depth_mean_batched(proto, coeff, det, depth)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); mean, count = masks.masked_depth_mean_batched(proto, coeff, det, depth); e1.record(); torch.cuda.synchronize()
print(f"mask+depth B={B}: {e0.elapsed_time(e1)*1e3:.1f} us; mean of means {mean[:, :160].nanmean().item():.4f} m")
