"""Role timeline of mask_umma_kernel (CTA 0): producer / MMA issuer / epilogue stamps per 128-pixel tile."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from types import SimpleNamespace
import tauv_vision_b200 as tv
from tauv_vision_b200.yolact.model import masks
lib = tv.load_library()
dev = torch.device("cuda", 0)
B, N, P, HP, TOPK, NK = 16, 19248, 32, 276, 200, 160
g = torch.Generator(device=dev); g.manual_seed(3)
coeff = torch.tanh(torch.randn((B, N, P), device=dev, generator=g))
proto = torch.nn.functional.leaky_relu(torch.randn((B, P, HP, HP), device=dev, generator=g))
keep = torch.randint(0, N, (B, TOPK), device=dev, generator=g)
n_keep = torch.full((B,), NK, dtype=torch.int32, device=dev)
box = torch.cat((torch.rand((B, TOPK, 2), device=dev, generator=g) * 0.8 + 0.1, torch.rand((B, TOPK, 2), device=dev, generator=g) * 0.4 + 0.05), -1)
det = SimpleNamespace(keep=keep, n_keep=n_keep, box=box)
out = torch.empty((B, TOPK, HP, HP), device=dev)
for _ in range(2): masks.assemble_mask_batched(proto, coeff, det, out=out)
torch.cuda.synchronize()
trace = torch.zeros((512, 8), dtype=torch.int64, device=dev)
lib.tauv_debug_mask_trace.argtypes = [ctypes.c_void_p]
lib.tauv_debug_mask_trace(trace.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); masks.assemble_mask_batched(proto, coeff, det, out=out); e1.record(); torch.cuda.synchronize()
lib.tauv_debug_mask_trace(None)
t = trace.cpu().numpy().astype(np.float64)
n = int((t[:, 5] > 0).sum())
t = t[:n]; t0 = t[0, 0] if t[0, 0] > 0 else t[:, 1].min()
u = (t - t0) / 1e3
print(f"kernel {e0.elapsed_time(e1)*1e3:.1f} us; CTA 0 handled {n} tiles -> {e0.elapsed_time(e1)*1e3/n:.2f} us per tile")
print("tile  prod:start  prod:done | mma:a_full  mma:acc_empty  mma:issued | epi:acc_full  epi:done")
for i in list(range(0, 8)) + list(range(30, 36)):
    if i < n: print(f"{i:4d} {u[i,0]:10.2f} {u[i,1]:10.2f} | {u[i,2]:10.2f} {u[i,3]:13.2f} {u[i,4]:11.2f} | {u[i,5]:12.2f} {u[i,6]:9.2f}")
m = slice(8, n)
print("means (us): producer work %.2f | producer wait for free stage %.2f | mma wait a_full->issue %.2f | epilogue wait %.2f | epilogue work %.2f | tile period %.2f" % (
    np.mean(t[m, 1] - t[m, 0]) / 1e3, np.mean(t[8:n, 0] - t[7:n-1, 1]) / 1e3, np.mean(t[m, 4] - t[m, 2]) / 1e3,
    np.mean(t[8:n, 5] - t[7:n-1, 6]) / 1e3, np.mean(t[m, 6] - t[m, 5]) / 1e3, (t[n-1, 6] - t[8, 6]) / 1e3 / (n - 9)))
