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
t = trace.cpu().numpy().astype(np.float64)[:n_items].reshape(B, C, 8)
t0 = t[:, :, 0][t[:, :, 0] > 0].min()
us = lambda x: (x - t0) / 1e3
first = t[:, :8, :]                       # the bootstrap items (rank r takes item r of the frame)
print("bootstrap items: tile stored at %.2f | tested+binned +%.2f | sync+threshold+filter +%.2f | first run starts +%.2f" % (
    us(first[:, :, 1]).mean(), (first[:, :, 6] - first[:, :, 1]).mean() / 1e3, (first[:, :, 7] - first[:, :, 6]).mean() / 1e3,
    (first[:, :, 0] - first[:, :, 7]).mean() / 1e3))
rounds = (t[:, :, 2] - t[:, :, 0]) / 1e3
print("rounds per item (us): first items %.2f | items 8-39 %.2f | items 40-79 %.2f" % (rounds[:, :8].mean(), rounds[:, 8:40].mean(), rounds[:, 40:].mean()))
# a run = items r, r+8, r+16, r+24 (then service); service duration = t3 - t2 of the run's last item
svc = (t[:, :, 3] - t[:, :, 2]) / 1e3
last_of_run = [j for j in range(C) if ((j // 8) % 4 == 3) or j // 8 == (C - 1) // 8]
print("service step (us): mean %.2f  max %.2f" % (svc[:, last_of_run].mean(), svc[:, last_of_run].max()))
for name, js in (("1st service", list(range(24, 32))), ("2nd service", list(range(56, 64))), ("last service", list(range(72, 80)))):
    r = t[:, js, :]
    print("  %s: wait for all warps +%.2f | peak tests (%d queued) +%.2f | bin+scan +%.2f | rest +%.2f" % (
        name, (r[:, :, 1] - r[:, :, 2]).mean() / 1e3, r[:, :, 4].mean(), (r[:, :, 6] - r[:, :, 1]).mean() / 1e3,
        (r[:, :, 7] - r[:, :, 6]).mean() / 1e3, (r[:, :, 3] - r[:, :, 7]).mean() / 1e3))
end = us(t[:, :, 3].max(axis=1))
print("unit (frame) end: min %.1f median %.1f max %.1f ; per-CTA end spread inside a cluster: mean %.2f us" % (
    end.min(), np.median(end), end.max(), np.mean([us(t[b, -8:, 3]).max() - us(t[b, -8:, 3]).min() for b in range(B)])))
thr = t[:, :, 5].astype(np.int64)
for j in (0, 8, 16, 32, 48, 72):
    print(f"  item {j:2d}: start {us(t[:, j, 0]).mean():6.1f}  rounds {rounds[:, j].mean():5.2f}  thr_key {int(np.median(thr[:, j])):#x}")
