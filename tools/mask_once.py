import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from types import SimpleNamespace
from tauv_vision_b200.yolact.model import masks
dev = torch.device("cuda", 0)
B, N, P, HP, TOPK = 16, 19248, 32, 276, 200
g = torch.Generator(device=dev); g.manual_seed(3)
coeff = torch.tanh(torch.randn((B, N, P), device=dev, generator=g))
proto = torch.nn.functional.leaky_relu(torch.randn((B, P, HP, HP), device=dev, generator=g))
keep = torch.randint(0, N, (B, TOPK), device=dev, generator=g)
n_keep = torch.full((B,), 160, dtype=torch.int32, device=dev)
box = torch.cat((torch.rand((B, TOPK, 2), device=dev, generator=g) * 0.8 + 0.1, torch.rand((B, TOPK, 2), device=dev, generator=g) * 0.4 + 0.05), -1)
det = SimpleNamespace(keep=keep, n_keep=n_keep, box=box)
out = torch.empty((B, TOPK, HP, HP), device=dev)
for _ in range(3):
    masks.assemble_mask_batched(proto, coeff, det, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); masks.assemble_mask_batched(proto, coeff, det, out=out); e1.record(); torch.cuda.synchronize()
byt = proto.numel()*4 + 160*B*HP*HP*4
print(f"mask B={B}: {e0.elapsed_time(e1)*1e3:.1f} us  {byt/e0.elapsed_time(e1)/1e6:.0f} GB/s")
