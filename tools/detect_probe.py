"""YOLACT detect at BASELINE.json configs[2], timed on the device with the host out of the picture (a spin kernel is
queued first, so every launch of the call is already enqueued when the first one starts), the L2 evicted by a READ
(no dirty lines left behind), and — in a -DTAUV_DEBUG build — the phase boundaries of frame 0's NMS CTA in cycles.
  TAUV_EXTRA_NVCC=-DTAUV_DEBUG python tools/detect_probe.py"""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import tauv_vision_b200 as tv
from tauv_vision_b200.yolact.model import nms

dev = torch.device("cuda", 0)
y = bench.make_yolact_inputs(dev, 3, 64)
lib = tv.load_library()
trace = None
if hasattr(lib, "tauv_debug_nms_trace"):
    trace = torch.zeros(16, dtype=torch.int64, device=dev)
    lib.tauv_debug_nms_trace.argtypes = [ctypes.c_void_p]
    lib.tauv_debug_nms_trace(ctypes.c_void_p(trace.data_ptr()))


def timeit(fn, n=9):
    ts = []
    for _ in range(n):
        y.proto.sum()                      # 624 MB read: evicts the L2 and leaves it clean
        torch.cuda._sleep(400000)          # ~200 us of spinning: the host runs ahead
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], r


res = {}
res["scores_us"], _ = timeit(lambda: nms.max_foreground_confidence(y.cls))
res["detect_us"], det = timeit(lambda: nms.detect(y.cls, y.enc, y.anchor, y.cfg, bench.YL_TOPK, bench.YL_IOU, bench.YL_CONF))
res["mean_n_keep"] = det.n_keep.float().mean().item()
res["detect_bytes"] = bench.algorithmic_bytes_yolact_detect(64, int(det.n_keep.sum()))
res["detect_gbs"] = res["detect_bytes"] / res["detect_us"] / 1e3
if trace is not None:
    t = trace.cpu().tolist()
    names = ["start", "scores loaded", "warp maxima sorted", "candidates listed", "candidates ranked", "boxes decoded",
             "pairs tested", "compacted"]
    res["nms_phase_cycles"] = {names[i]: t[i] - t[i - 1] for i in range(1, 8)}
    res["nms_phase_cycles"]["classes"] = t[9] - t[7]
    res["nms_total_cycles"] = t[9] - t[0]
    res["n_candidates_frame0"] = t[8]
print(json.dumps(res, indent=1))
