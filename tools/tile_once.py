"""ncu target: four runs of stage 1 (seed + tile kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tauv_vision_b200 as tv
from tauv_vision_b200 import _lib
lib = tv.load_library()
dev = torch.device("cuda", 0)
B, C, H, W, K = 64, 80, 128, 128, 100
g = torch.Generator(device=dev); g.manual_seed(1)
logits = torch.randn((B, C, H, W), device=dev, generator=g) * 1.5 - 2.2
ws = torch.empty(lib.tauv_heatmap_topk_workspace_bytes(B, C, H, W, K), dtype=torch.uint8, device=dev)
def run():
    rc = lib.tauv_heatmap_topk_stage1(_lib.fptr(logits), B, C, H, W, K, 1, ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
    assert rc == 0
    torch.cuda.synchronize()
run(); run(); run(); run()
print("done")
