"""Role timeline of mask_umma_kernel (CTA 0): producer / MMA issuer / epilogue stamps per 128-pixel tile."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from types import SimpleNamespace
import tauv_vision_b200 as tv
from tauv_vision_b200.yolact.model import masks
lib = tv.load_library()
dev = torch.device("cuda", 0)
B, N, P, HP, TOPK, NK = int(os.environ.get("B", 16)), 19248, 32, 276, 200, 160
g = torch.Generator(device=dev); g.manual_seed(3)
coeff = torch.tanh(torch.randn((B, N, P), device=dev, generator=g))
proto = torch.nn.functional.leaky_relu(torch.randn((B, P, HP, HP), device=dev, generator=g))
keep = torch.randint(0, N, (B, TOPK), device=dev, generator=g)
n_keep = torch.full((B,), NK, dtype=torch.int32, device=dev)
box = torch.cat((torch.rand((B, TOPK, 2), device=dev, generator=g) * 0.8 + 0.1, torch.rand((B, TOPK, 2), device=dev, generator=g) * 0.4 + 0.05), -1)
det = SimpleNamespace(keep=keep, n_keep=n_keep, box=box)
out = torch.empty((B, TOPK, HP, HP), device=dev)
if os.environ.get("DEPTH") == "1":  # the fused consumer (masked depth mean) instead of the mask writer
    depth = torch.randint(300, 9000, (B, 720, 1280), device=dev, dtype=torch.int32).to(torch.uint16)
    ws = torch.empty(lib.tauv_yolact_mask_depth_workspace_bytes(B, HP, HP, TOPK), dtype=torch.uint8, device=dev)
    run = lambda: masks.masked_depth_mean_batched(proto, coeff, det, depth, workspace=ws)
else:
    run = lambda: masks.assemble_mask_batched(proto, coeff, det, out=out)
for _ in range(2): run()
torch.cuda.synchronize()
trace = torch.zeros((512, 8), dtype=torch.int64, device=dev)
lib.tauv_debug_mask_trace.argtypes = [ctypes.c_void_p]
lib.tauv_debug_mask_trace(trace.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
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

if os.environ.get("DEPTH") == "1":
    raw = trace.cpu().numpy()[:n, 7].astype(np.uint64)
    slow_t = (raw >> np.uint64(5)).astype(np.float64); slow_w = (raw & np.uint64(31)).astype(int)
    t6 = (trace.cpu().numpy()[:n, 6].astype(np.uint64) & np.uint64((1 << 59) - 1)).astype(np.float64)  # (the shift dropped the top bits)
    lag = (slow_t - t6) / 1e3
    print("slowest epilogue warp per tile: lag behind warp 0 mean %.2f us max %.2f; which warp: %s" % (lag[8:].mean(), lag[8:].max(), np.bincount(slow_w[8:], minlength=12).tolist()))
    if os.environ.get("DBG3") == "1":  # built with -DTAUV_DEPTH_DBG=3: columns 2, 3 hold cycle counts of warp 0's epilogue
        raw = trace.cpu().numpy()[:n]
        print("warp 0 epilogue cycles per tile: tcgen05.ld + wait %.0f, whole chunk loop %.0f" % (raw[8:, 2].mean(), raw[8:, 3].mean()))
