"""Latency of the YOLACT node sequence at batch 1: box_decode -> nms -> assemble_mask (yolact_node.py:127-130), and of the
fused detect + assemble_mask_batched."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tauv_vision_b200.yolact.model import nms, masks, boxes, anchors
from tests import synth
dev = torch.device("cuda", 0)
cfg = synth.yolact_config()
B, N, C1, P, HP, TOPK = 1, 19248, 81, 32, 276, 200
g = torch.Generator(device=dev); g.manual_seed(3)
anchor = anchors.all_anchors(synth.fpn_sizes(550, 550), cfg, dev)
cls = torch.randn((B, N, C1), device=dev, generator=g) * 2
cls[:, :, 0] += 4
idx = torch.randint(0, N - 16, (B, 12), device=dev, generator=g)
for j in range(12):
    for o in range(12):
        cls[torch.arange(B, device=dev), idx[:, j] + o, 1 + (j % (C1 - 1))] += 10 + torch.rand((B,), device=dev, generator=g) * 4
enc = torch.randn((B, N, 4), device=dev, generator=g) * 0.3
coeff = torch.tanh(torch.randn((B, N, P), device=dev, generator=g))
proto = torch.nn.functional.leaky_relu(torch.randn((B, P, HP, HP), device=dev, generator=g))

def node_sequence():
    box = boxes.box_decode(enc, anchor, cfg)
    keep = nms.nms(cls, box, TOPK, 0.5, 0.05)
    m = masks.assemble_mask(proto[0], coeff[0, keep], box[0, keep])
    return keep, m

def fused():
    det = nms.detect(cls, enc, anchor, cfg, TOPK, 0.5, 0.05)
    return det, masks.assemble_mask_batched(proto, coeff, det)

for name, fn in (("node sequence (box_decode, nms, assemble_mask)", node_sequence), ("fused (detect, assemble_mask_batched)", fused)):
    for _ in range(3): r = fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20): r = fn(); torch.cuda.synchronize()
    t1 = time.perf_counter()
    n = r[0].numel() if torch.is_tensor(r[0]) else int(r[0].n_keep[0])
    print(f"B=1 {name}: {1e6*(t1-t0)/20:8.1f} us per frame ({n} kept)")
