"""Time the YOLACT post-process kernels at BASELINE.json configs[2]: B=64, N=19248, 81 classes, top_k=200, 32 protos 276x276."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from types import SimpleNamespace
import tauv_vision_b200 as tv
from tauv_vision_b200.yolact.model import nms, masks, boxes, anchors, loss
from tests import synth

dev = torch.device("cuda", 0)
B, N, C1, P, HP, TOPK = 64, 19248, 81, 32, 276, 200
if len(sys.argv) > 1: B = int(sys.argv[1])
cfg = synth.yolact_config()
g = torch.Generator(device=dev); g.manual_seed(3)
anchor = anchors.all_anchors(synth.fpn_sizes(550, 550), cfg, dev)
cls = torch.randn((B, N, C1), device=dev, generator=g) * 2
cls[:, :, 0] += 4
# planted confident clusters: ~150 priors per frame in overlapping groups
idx = torch.randint(0, N - 16, (B, 12), device=dev, generator=g)
for j in range(12):
    for o in range(12):
        cls[torch.arange(B, device=dev), idx[:, j] + o, 1 + (j % (C1 - 1))] += 10 + torch.rand((B,), device=dev, generator=g) * 4
enc = torch.randn((B, N, 4), device=dev, generator=g) * 0.3
coeff = torch.tanh(torch.randn((B, N, P), device=dev, generator=g))
proto = torch.nn.functional.leaky_relu(torch.randn((B, P, HP, HP), device=dev, generator=g))
out = torch.empty((B, TOPK, HP, HP), device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, n=10):
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], r

res = {}
t, score = timeit(lambda: nms.max_foreground_confidence(cls))
res["scores_us"] = t; res["scores_gbs"] = cls.numel() * 4 / t / 1e3
t, det = timeit(lambda: nms.detect(cls, enc, anchor, cfg, TOPK, 0.5, 0.05))
res["detect_us"] = t; res["detect_gbs"] = (cls.numel() * 4 + 2 * B * N * 16) / t / 1e3
nk = det.n_keep.float().mean().item(); res["mean_n_keep"] = nk
t, _ = timeit(lambda: masks.assemble_mask_batched(proto, coeff, det, out=out), n=5)
mask_bytes = proto.numel() * 4 + det.n_keep.sum().item() * (HP * HP * 4 + P * 4 + 16)
mask_flops = 2.0 * det.n_keep.sum().item() * P * HP * HP
res["mask_us"] = t; res["mask_gbs"] = mask_bytes / t / 1e3; res["mask_tflops_useful"] = mask_flops / t / 1e6
res["mask_tflops_issued"] = 3 * 2.0 * B * ((HP * HP + 255) // 256) * 128 * 256 * P / t / 1e6
# fused consumer: mask assembly + nearest resize to the camera + mean depth under each mask (no mask written)
HI, WI = 720, 1280
depth = torch.randint(300, 9000, (B, HI, WI), device=dev, dtype=torch.int32)
depth = torch.where(torch.rand((B, HI, WI), device=dev) < 0.2, torch.zeros((), dtype=torch.int32, device=dev), depth).to(torch.uint16)
ws_d = torch.empty(tv.load_library().tauv_yolact_mask_depth_workspace_bytes(B, HP, HP, TOPK), dtype=torch.uint8, device=dev)
t, (dm, dc) = timeit(lambda: masks.masked_depth_mean_batched(proto, coeff, det, depth, workspace=ws_d), n=5)
depth_bytes = proto.numel() * 4 + depth.numel() * 2 + 2 * B * HP * HP * 8 + det.n_keep.sum().item() * (P * 4 + 16 + 16)
res["mask_depth_us"] = t; res["mask_depth_gbs"] = depth_bytes / t / 1e3
res["mask_depth_camera"] = [HI, WI]; res["mask_depth_valid_rows"] = int((dc > 0).sum().item())
res["frames_per_s_detect_plus_mask_depth"] = B / ((res["detect_us"] + t) * 1e-6)
os.environ["TAUV_MASK_SIMT"] = "1"
t, _ = timeit(lambda: masks.assemble_mask_batched(proto, coeff, det, out=out), n=3)
del os.environ["TAUV_MASK_SIMT"]
res["mask_simt_us"] = t
t, _ = timeit(lambda: boxes.box_decode(enc, anchor, cfg)); res["box_decode_us"] = t; res["box_decode_gbs"] = 2 * B * N * 16 / t / 1e3
tb, tvd = synth.truth_boxes(B, 16, seed=1)
tb, tvd = tb.to(dev), tvd.to(dev)
t, _ = timeit(lambda: loss.match_anchors(anchor, tb, tvd, cfg)); res["match_us"] = t
res["match_gbs"] = B * N * 30 / t / 1e3
res["frames_per_s_full_postprocess"] = B / ((res["detect_us"] + res["mask_us"]) * 1e-6)
print(json.dumps(res, indent=1))
