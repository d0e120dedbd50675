"""Timing target: the three heads' level outputs packed into [B,N,C] (B = 64, five levels of 550 x 550, 81 classes, 4 box
values, 32 coefficients) against the reference's permute + reshape (+ tanh) + cat as eager torch ops on the same GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tauv_vision_b200.yolact.model import prediction_head as PH
from tests import synth
dev = torch.device("cuda", 0)
B = 64
sizes = synth.fpn_sizes(550, 550)
g = torch.Generator(device=dev); g.manual_seed(1)
lv = {C: [torch.randn((B, 3 * C, h, w), device=dev, generator=g) for h, w in sizes] for C in (81, 4, 32)}

def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); torch.cuda._sleep(2_000_000)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

def eager(C, th):
    out = torch.cat([t.permute(0, 2, 3, 1).reshape(B, -1, C) for t in lv[C]], dim=1)
    return torch.tanh(out) if th else out
for C, th in ((81, False), (4, False), (32, True)):
    nb = 2 * 4 * B * 19248 * C
    t_ours = timed(lambda: PH.pack_head(lv[C], C, tanh=th))
    t_ref = timed(lambda: eager(C, th))
    print(f"pack_head C={C}{' + tanh' if th else ''}: {t_ours:.1f} us = {nb / t_ours / 1e3:.0f} GB/s (read + write once); "
          f"eager permute/reshape/cat{'/tanh' if th else ''}: {t_ref:.1f} us")
