"""Phase timeline of select_kernel (debug hook tauv_debug_select_trace; needs a -DTAUV_DEBUG build:
    TAUV_EXTRA_NVCC=-DTAUV_DEBUG python tools/select_trace.py [B C H W K])."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import torch, numpy as np
import tauv_vision_b200 as tv
from tauv_vision_b200.centernet.model import decode as D
lib = tv.load_library()
dev = torch.device("cuda", 0)
a = [int(x) for x in sys.argv[1:]]
B, C, H, W, K = (a + [64, 80, 128, 128, 100][len(a):])
g = torch.Generator(device=dev); g.manual_seed(1)
logits = torch.randn((B, C, H, W), device=dev, generator=g) * 1.5 - 2.2
if os.environ.get("SMOOTH"):  # box-filtered noise (SMOOTH="box,passes"): the smooth maps of a trained head
    box, passes = (int(v) for v in os.environ["SMOOTH"].split(","))
    x = torch.randn((B, C, H + passes * (box - 1), W + passes * (box - 1)), device=dev, generator=g)
    for _ in range(passes):
        x = torch.nn.functional.avg_pool2d(x, box, 1)
    logits = ((x - x.mean()) / x.std() * 1.5 - 2.2).contiguous()
size = (torch.rand((B, 2, H, W), device=dev, generator=g) * 0.3).permute(0, 2, 3, 1)
offset = (torch.rand((B, 2, H, W), device=dev, generator=g) * 4).permute(0, 2, 3, 1)
mc = SimpleNamespace(in_h=H * 4, in_w=W * 4, downsample_ratio=4, out_h=H, out_w=W)
pred = SimpleNamespace(heatmap=logits, size=size, offset=offset, depth=None)
out = D.decode_packed(pred, mc, K, 0.3)
trace = torch.zeros((B, 32), dtype=torch.int64, device=dev)
lib.tauv_debug_select_trace.argtypes = [ctypes.c_void_p]
for _ in range(3):
    D.decode_packed(pred, mc, K, 0.3, out=out)
torch.cuda.synchronize()
lib.tauv_debug_select_trace(trace.data_ptr())
D.decode_packed(pred, mc, K, 0.3, out=out)
torch.cuda.synchronize()
lib.tauv_debug_select_trace(None)
tt = trace.cpu().numpy().astype(np.float64)
for rep in range(2):
  t = tt[:, rep * 16:(rep + 1) * 16]
  if not (t[:, 1] > 0).any():
    continue
  print("pass", rep + 1, "(the second pass exists only in -DTAUV_SEL_TWICE builds: same work, warm caches)")
  t0 = t[:, 1].min()
  names = {0: "start", 1: "dep wait done", 2: "threshold", 3: "hot blocks", 11: "  rows loaded", 12: "  neighbours", 4: "examined",
           13: "  ranked", 5: "ranked+emitted", 6: "finished"}
  print(f"select_kernel phases, B={B} C={C} {H}x{W} k={K} (us after the first CTA passed griddepcontrol.wait; mean / max over frames)")
  for i, nme in names.items():
      col = t[:, i][t[:, i] > 0]
      if len(col):
          print(f"  {nme:16s} {((col - t0) / 1e3).mean():7.2f} {((col - t0) / 1e3).max():7.2f}")
  print("  attempts per frame:", np.unique(t[:, 8], return_counts=True), " hot blocks mean %.1f max %d, queued cells mean %.1f min %d" % (
      t[:, 9].mean(), t[:, 9].max(), t[:, 10].mean(), t[:, 10].min()), " peaks in the list mean %.1f min %d" % (t[:, 14].mean(), t[:, 14].min()))
  print("  CTA launched before the dependency resolved by (us): mean %.2f max %.2f" % (((t[:, 1] - t[:, 0]) / 1e3).mean(), ((t[:, 1] - t[:, 0]) / 1e3).max()))
