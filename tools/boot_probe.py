"""Is the bootstrap's cost a cold-start effect?  Two units per cluster (B = 2 x resident clusters): compare the phases of
the first and the second unit of each cluster."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import tauv_vision_b200 as tv
from tauv_vision_b200 import _lib
lib = tv.load_library()
dev = torch.device("cuda", 0)
B, C, H, W, K = 142, 80, 128, 128, 100
g = torch.Generator(device=dev); g.manual_seed(1)
logits = torch.randn((B, C, H, W), device=dev, generator=g) * 1.5 - 2.2
ws = torch.empty(lib.tauv_heatmap_topk_workspace_bytes(B, C, H, W, K), dtype=torch.uint8, device=dev)
trace = torch.zeros((B * C, 8), dtype=torch.int64, device=dev)
def run():
    rc = lib.tauv_heatmap_topk_stage1(_lib.fptr(logits), B, C, H, W, K, 1, ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
    assert rc == 0; torch.cuda.synchronize()
run(); run()
lib.tauv_debug_tile_trace.argtypes = [ctypes.c_void_p]
lib.tauv_debug_tile_trace(trace.data_ptr()); run(); lib.tauv_debug_tile_trace(None)
t = trace.cpu().numpy().astype(np.float64).reshape(B, C, 8)
first = t[:, :8, :]
for name, sl in (("first unit of a cluster ", slice(0, 71)), ("second unit of a cluster", slice(71, 142))):
    f = first[sl]
    print("%s: tests +%.2f | barrier +%.2f | binning +%.2f | sync+threshold+filter +%.2f | to stream start +%.2f" % (
        name, (f[:, :, 3] - f[:, :, 1]).mean() / 1e3, (f[:, :, 4] - f[:, :, 3]).mean() / 1e3, (f[:, :, 6] - f[:, :, 4]).mean() / 1e3,
        (f[:, :, 7] - f[:, :, 6]).mean() / 1e3, (f[:, :, 0] - f[:, :, 7]).mean() / 1e3))
