"""Experiment driver: time stage 1 (tile_topk_kernel) alone under different conditions."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tauv_vision_b200 as tv
from tauv_vision_b200 import _lib

lib = tv.load_library()
dev = torch.device("cuda", 0)
B, C, H, W, K = 64, 80, 128, 128, 100
if len(sys.argv) > 1:
    B, C, H, W, K = map(int, sys.argv[1:6])
g = torch.Generator(device=dev); g.manual_seed(1)
logits = torch.randn((B, C, H, W), device=dev, generator=g) * 1.5 - 2.2
nbytes = lib.tauv_heatmap_topk_workspace_bytes(B, C, H, W, K)
ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def run(n=20, label=""):
    ts = []
    for i in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.tauv_heatmap_topk_stage1(_lib.fptr(logits), B, C, H, W, K, 1, ws.data_ptr(), ws.numel(),
                                          _lib.stream_ptr(dev))
        e1.record()
        assert rc == 0, lib.tauv_last_error()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    gb = 4 * B * C * H * W / 1e9
    print(f"{label:28s} median {ts[len(ts)//2]:8.1f} us  min {ts[0]:8.1f} us  -> {gb / (ts[len(ts)//2] * 1e-6):7.0f} GB/s")

run(label="seed + stream kernel")
# copy bandwidth reference
a = torch.empty(335544320 // 4, device=dev); b = torch.empty_like(a)
for _ in range(3): b.copy_(a)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); b.copy_(a); e1.record(); torch.cuda.synchronize()
print(f"torch copy 335 MB: {e0.elapsed_time(e1)*1e3:.1f} us -> {2*0.3355/(e0.elapsed_time(e1)*1e-3):.0f} GB/s (r+w)")
s = torch.empty((), device=dev)
e0.record(); s = logits.max(); e1.record(); torch.cuda.synchronize()
e0.record(); s = logits.max(); e1.record(); torch.cuda.synchronize()
print(f"torch max-reduce 335 MB read: {e0.elapsed_time(e1)*1e3:.1f} us -> {0.3355/(e0.elapsed_time(e1)*1e-3):.0f} GB/s (read only)")
