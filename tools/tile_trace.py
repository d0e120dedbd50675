"""Per-item timeline of tile_topk_kernel (debug hook tauv_debug_tile_trace)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import tauv_vision_b200 as tv
from tauv_vision_b200 import _lib
lib = tv.load_library()
dev = torch.device("cuda", 0)
B, C, H, W, K = 64, 80, 128, 128, 100
g = torch.Generator(device=dev); g.manual_seed(1)
logits = torch.randn((B, C, H, W), device=dev, generator=g) * 1.5 - 2.2
ws = torch.empty(lib.tauv_heatmap_topk_workspace_bytes(B, C, H, W, K), dtype=torch.uint8, device=dev)
n_items = B * C
trace = torch.zeros((n_items, 8), dtype=torch.int64, device=dev)
def run():
    rc = lib.tauv_heatmap_topk_stage1(_lib.fptr(logits), B, C, H, W, K, 1, ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
    assert rc == 0; torch.cuda.synchronize()
run(); run()
lib.tauv_debug_tile_trace.argtypes = [ctypes.c_void_p]
lib.tauv_debug_tile_trace(trace.data_ptr())
run()
lib.tauv_debug_tile_trace(None)
t = trace.cpu().numpy().astype(np.float64)
t0 = t[:, 0].min()
start, boot, scan, end = [(t[:, i] - t0) / 1e3 for i in range(4)]
print(f"kernel span {end.max():.1f} us; items {n_items}")
print("blk   start    boot_dur scan_dur  fin_dur  n_list  thr_key")
for b in list(range(0, 40, 4)) + list(range(560, 700, 20)) + list(range(1000, 5120, 400)):
    print(f"{b:5d} {start[b]:8.1f} {boot[b]-start[b]:8.2f} {scan[b]-boot[b]:8.2f} {end[b]-scan[b]:8.2f} {int(t[b,4]):6d}  {int(t[b,5]):#x}")
for lo, hi in [(0, 592), (592, 1184), (1184, 2368), (2368, 5120)]:
    s = slice(lo, hi)
    print(f"blocks [{lo},{hi}): start {start[s].min():6.1f}-{start[s].max():6.1f}  mean dur {np.mean(end[s]-start[s]):6.2f} "
          f"(boot {np.mean(boot[s]-start[s]):5.2f} scan {np.mean(scan[s]-boot[s]):5.2f} fin {np.mean(end[s]-scan[s]):5.2f})  "
          f"mean list {t[s,4].mean():7.1f}  no-thr {int((t[s,5]==0).sum())}")
