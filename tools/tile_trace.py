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
print("bootstrap detail: peak tests +%.2f | barrier +%.2f | binning etc +%.2f" % ((first[:, :, 3] - first[:, :, 1]).mean() / 1e3, (first[:, :, 4] - first[:, :, 3]).mean() / 1e3, (first[:, :, 6] - first[:, :, 4]).mean() / 1e3))
last = t[:, -8:, :]
print("kernel start (CTA's first instruction) at %.2f .. %.2f; CTA end at %.1f .. %.1f" % (us(last[:, :, 6]).min(), us(last[:, :, 6]).max(), us(last[:, :, 3]).min(), us(last[:, :, 3]).max()))
smid = last[:, :, 7].astype(int).ravel(); cend = us(last[:, :, 2]).ravel()
per_sm = np.bincount(smid, minlength=148)
print("CTAs per SM histogram:", np.bincount(per_sm), " mean stream end by CTAs-on-SM:", {int(c): round(float(cend[per_sm[smid] == c].mean()), 1) for c in np.unique(per_sm[smid])})
rounds = (t[:, 8:, 2] - t[:, 8:, 0]) / 1e3
rounds = np.concatenate([np.zeros((B, 8)), rounds], axis=1)
print("rounds per item (us): items 8-39 %.2f | items 40-79 %.2f" % (rounds[:, 8:40].mean(), rounds[:, 40:].mean()))
end = us(t[:, :, 3].max(axis=1))
se = us(t[:, -8:, 2])  # end of streaming of each CTA (its last item)
print("unit (frame) end: min %.1f median %.1f max %.1f ; streaming end per CTA: min %.1f max %.1f, spread inside a cluster mean %.2f us; drain after streaming mean %.2f us; list at end mean %.1f" % (
    end.min(), np.median(end), end.max(), se.min(), se.max(), np.mean(se.max(axis=1) - se.min(axis=1)), np.mean(us(t[:, -8:, 3]) - se), t[:, -8:, 4].mean()))
thr = t[:, :, 5].astype(np.int64)
for j in (8, 16, 32, 48, 72):
    print(f"  item {j:2d}: start {us(t[:, j, 0]).mean():6.1f}  rounds {rounds[:, j].mean():5.2f}  thr_key {int(np.median(thr[:, j])):#x}")
