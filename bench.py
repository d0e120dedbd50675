#!/usr/bin/env python
"""bench.py — decoded frames/s of the detection-head hot path on N B200s, one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload centernet|mixed]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workloads
  centernet (default; BASELINE.json configs[1], the line's `value`): CenterNet decode + Gaussian target encode of a
      batch of 64 synthetic 512x512 frames per GPU (stride 4 -> 80 classes x 128x128 heatmaps, top-100, 16 objects per
      frame).  One step = decode (1 kernel) + encode (1 kernel).  Weak scaling: every GPU takes its own 64 frames, no
      collective on the data path (NCCL only carries the timing scalars).  In the same run, on every rank, the YOLACT
      post-process kernels of configs[2] (B = 64, 19 248 priors, 81 classes, top_k 200, 32 prototypes 276x276) are
      timed as well and reported under `kernels` with their algorithmic bytes.
  mixed (configs[3]): 256 frames, each with a CenterNet head and a YOLACT head, sharded over the N GPUs
      (`shard.frame_range`, strong scaling).  One step = CenterNet decode + YOLACT detect + fused mask/depth consumer of
      the rank's frames; `e2e` additionally copies the packed results to the host and gathers them on rank 0 in frame
      order (`shard.gather_frames`) inside the timed region — the "final host gather" of SURVEY.md section 8e.

value     : frames/s with inputs resident in HBM; the K-step block is timed `--blocks` (>= 5) times with CUDA events on
            the launching stream and the MEDIAN block is reported (`block_ms` lists them all); max over ranks.
e2e       : the same step through the public Python API with pinned HOST inputs copied in (ONE staging buffer per
            step and direction) and the packed results copied out inside the timed region, double-buffered over two
            streams; PCIe-bound.  Each rank pins itself and its staging memory to its GPU's NUMA node; the per-rank
            host->device rate is reported (`h2d_gbs_per_rank`).
roofline  : the whole decode call — block_max_kernel (the only pass over the 335.5 MB of logits; the dominant kernel) +
            select_kernel (per-frame select on the 3 % summaries, programmatically dependent) — timed live with events
            around the call inside the timed steps, against MEASURED_PEAKS.json; `frac` is the WHOLE call's algorithmic
            bytes over its time (the harder number, comparable with round 1's single kernel).  `dominant_kernel` is
            block_max_kernel alone (tauv_centernet_block_maxima, events, back to back).  In the steps the decode runs right
            after the target encode, whose dirty L2 lines are written back while it streams
            (profiles/r2_stream_bench_v*.txt); `us_per_launch_isolated` is the same call timed back to back with itself.
cpu_baseline / --impl reference : the CPU oracle port (oracle/ref_port.py, torch-CPU with all host threads) on the
            FULL 64-frame batch per step.  (It times the tensor part of decode without the reference's per-detection
            Python loop, decode.py:204-234, which flatters the CPU.)
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path
from types import SimpleNamespace

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

B_PER_GPU, C, H, W, K_DET, N_OBJ = 64, 80, 128, 128, 100, 16
IN_HW, DOWNSAMPLES, SIGMA, THR = 512, 2, 2.0, 0.3
YL_N, YL_C1, YL_P, YL_HP, YL_TOPK, YL_IOU, YL_CONF = 19248, 81, 32, 276, 200, 0.5, 0.05
CAM_H, CAM_W = 720, 1280
MIXED_FRAMES = 256
METRIC, UNIT = "decoded_frames_per_sec", "frames/s"
WORKLOAD = "centernet_decode_topk100+gaussian_target_encode, batch 64 x [80,128,128] per GPU (BASELINE configs[1], stride 4)"
WORKLOAD_MIXED = ("mixed CenterNet (80x128x128, top-100) + YOLACT (19248 priors, 81 classes, top_k 200, 32 protos 276x276, "
                  "depth 720x1280) head decode, 256 frames sharded over the GPUs (BASELINE configs[3])")


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0))), "measured"
    return 6650.0, 1400.0, "fallback"


def algorithmic_bytes_decode(B):
    # read the logits once + index/label/score (2*8+8+4) + size/offset gathers (16) + packed boxes (24)
    return 4 * B * C * H * W + B * K_DET * (28 + 16 + 24)


def algorithmic_bytes_encode(B):
    return 4 * B * C * H * W


def algorithmic_bytes_yolact_detect(B, n_keep_total):
    # class logits once + box encodings + anchors once per batch + kept indices
    return B * (4 * YL_N * YL_C1 + 16 * YL_N) + 16 * YL_N + 8 * n_keep_total


def algorithmic_bytes_mask(B, n_keep_total):
    # prototypes once + per kept mask: coefficients, box, the fp32 mask itself
    return B * 4 * YL_P * YL_HP * YL_HP + n_keep_total * (YL_P * 4 + 16 + 4 * YL_HP * YL_HP)


def algorithmic_bytes_mask_depth(B, n_keep_total):
    # prototypes once + the mono16 depth image once + per kept mask: coefficients, box, (mean, count)
    return B * (4 * YL_P * YL_HP * YL_HP + 2 * CAM_H * CAM_W) + n_keep_total * (YL_P * 4 + 16 + 16)


def algorithmic_bytes_mask_binary(B, n_keep_total):
    # prototypes once + per kept mask: coefficients, box, ONE byte per camera pixel (the binarised upsampled mask)
    return B * 4 * YL_P * YL_HP * YL_HP + n_keep_total * (YL_P * 4 + 16 + CAM_H * CAM_W)


class ClockSampler:
    """nvidia-smi-equivalent clock / throttle-reason sampling through NVML during the timed region."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t is not None:
            self._t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def pin_to_gpu_numa_node(device_index: int):
    """Pin this process (and, by first touch, the pinned staging memory it allocates afterwards) to the cores of the
    GPU's NUMA node.  Returns (node, n_cores) or (None, n_cores) when the topology cannot be read."""
    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = Path(f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node")
        node = int(path.read_text().strip())
        if node < 0:
            return None, len(os.sched_getaffinity(0))
        cpus = Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip()
        cores = set()
        for part in cpus.split(","):
            lo, _, hi = part.partition("-")
            cores.update(range(int(lo), int(hi or lo) + 1))
        cores &= os.sched_getaffinity(0)
        if cores:
            os.sched_setaffinity(0, cores)
        return node, len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        return None, len(os.sched_getaffinity(0))


# ---------------------------------------------------------------------------------------------------------------------
# synthetic inputs
# ---------------------------------------------------------------------------------------------------------------------
def make_inputs(device, seed, B=B_PER_GPU):
    """Synthetic CenterNet head tensors generated on the device (N(-2.2,1.5) logits = the reference's heatmap bias init)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    logits = torch.randn((B, C, H, W), device=device, generator=g) * 1.5 - 2.2
    size = (torch.rand((B, 2, H, W), device=device, generator=g) * 0.3).permute(0, 2, 3, 1)
    offset = (torch.rand((B, 2, H, W), device=device, generator=g) * 4).permute(0, 2, 3, 1)
    truth = SimpleNamespace(valid=torch.rand((B, N_OBJ), device=device, generator=g) < 0.75,
                            label=torch.randint(0, C, (B, N_OBJ), device=device, generator=g),
                            center=torch.rand((B, N_OBJ, 2), device=device, generator=g))
    return logits, size, offset, truth


def make_yolact_inputs(device, seed, B):
    """Synthetic YOLACT head tensors (configs[2]): N(0,2) class logits with a background bias and ~150 planted confident
    priors per frame in overlapping clusters (so NMS keeps ~100-160), N(0,0.3) box encodings, tanh coefficients,
    leaky-relu prototypes, mono16 depth with 20 % holes; anchors from get_anchor at 550x550."""
    from tauv_vision_b200.yolact.model import anchors
    from tests import synth
    cfg = synth.yolact_config()
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    anchor = anchors.all_anchors(synth.fpn_sizes(550, 550), cfg, device)
    cls = torch.randn((B, YL_N, YL_C1), device=device, generator=g) * 2
    cls[:, :, 0] += 4
    idx = torch.randint(0, YL_N - 16, (B, 12), device=device, generator=g)
    rows = torch.arange(B, device=device)
    for j in range(12):
        for o in range(12):
            cls[rows, idx[:, j] + o, 1 + (j % (YL_C1 - 1))] += 10 + torch.rand((B,), device=device, generator=g) * 4
    enc = torch.randn((B, YL_N, 4), device=device, generator=g) * 0.3
    coeff = torch.tanh(torch.randn((B, YL_N, YL_P), device=device, generator=g))
    proto = torch.nn.functional.leaky_relu(torch.randn((B, YL_P, YL_HP, YL_HP), device=device, generator=g))
    depth = torch.randint(300, 9000, (B, CAM_H, CAM_W), device=device, dtype=torch.int32, generator=g)
    holes = torch.rand((B, CAM_H, CAM_W), device=device, generator=g) < 0.2
    depth = torch.where(holes, torch.zeros((), dtype=torch.int32, device=device), depth).to(torch.uint16)
    return SimpleNamespace(cfg=cfg, anchor=anchor, cls=cls, enc=enc, coeff=coeff, proto=proto, depth=depth)


# ---------------------------------------------------------------------------------------------------------------------
# the CPU arm
# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_sample(n_frames: int, reps: int, warmup: int = 1):
    """The oracle port on the host cores: decode + encode of n_frames frames, `reps` times; returns frames/s."""
    from oracle import ref_port as O
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(1)
    logits = torch.randn((n_frames, C, H, W), generator=g) * 1.5 - 2.2
    size = (torch.rand((n_frames, 2, H, W), generator=g) * 0.3).permute(0, 2, 3, 1)
    offset = (torch.rand((n_frames, 2, H, W), generator=g) * 4).permute(0, 2, 3, 1)
    valid = torch.rand((n_frames, N_OBJ), generator=g) < 0.75
    label = torch.randint(0, C, (n_frames, N_OBJ), generator=g)
    center = torch.rand((n_frames, N_OBJ, 2), generator=g)

    def step():
        O.decode_packed(logits, size, offset, None, 2 ** DOWNSAMPLES, IN_HW, IN_HW, K_DET, THR, canonical=False)
        O.generate_heatmap(valid, label, center, C, H, W, IN_HW, IN_HW, 2 ** DOWNSAMPLES, SIGMA)

    for _ in range(max(1, warmup)):
        step()
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    dt = time.perf_counter() - t0
    return n_frames * reps / dt, dt / reps


def cpu_reference_mixed(n_frames: int, reps: int):
    """The oracle port of one mixed step on a bounded sample of frames: CenterNet decode of the sample + per frame
    box_decode -> nms -> assemble_mask -> nearest resize -> masked depth mean (yolact_node.py:127-135,178)."""
    from oracle import ref_port as O
    from tests import synth
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(2)
    logits = torch.randn((n_frames, C, H, W), generator=g) * 1.5 - 2.2
    size = (torch.rand((n_frames, 2, H, W), generator=g) * 0.3).permute(0, 2, 3, 1)
    offset = (torch.rand((n_frames, 2, H, W), generator=g) * 4).permute(0, 2, 3, 1)
    cfg = synth.yolact_config()
    anchor = torch.cat([O.get_anchor(i, s, cfg.anchor_scales, cfg.anchor_aspect_ratios, cfg.in_h, cfg.in_w)
                        for i, s in enumerate(synth.fpn_sizes(550, 550))], dim=1)
    cls, enc = synth.yolact_heads(n_frames, YL_N, YL_C1, seed=5, anchor=anchor)
    coeff = torch.tanh(torch.randn((n_frames, YL_N, YL_P), generator=g))
    proto = torch.nn.functional.leaky_relu(torch.randn((n_frames, YL_P, YL_HP, YL_HP), generator=g))
    depth = torch.stack([synth.depth_image(CAM_H, CAM_W, seed=7 + i) for i in range(n_frames)])

    def step():
        O.decode_packed(logits, size, offset, None, 2 ** DOWNSAMPLES, IN_HW, IN_HW, K_DET, THR, canonical=False)
        for b in range(n_frames):
            box = O.box_decode(enc[b:b + 1], anchor, cfg.box_variances)
            keep = O.nms(cls[b:b + 1], box, YL_TOPK, YL_IOU, YL_CONF)
            if keep.numel():
                O.masked_depth_mean(proto[b], coeff[b, keep], box[0, keep], depth[b])

    step()
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    dt = time.perf_counter() - t0
    return n_frames * reps / dt, dt / reps


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; the reference is pure
    Python/torch and cannot travel to the GPU box) on the host cores, on the FULL batch per step.  Rank 0 only."""
    if rank != 0:
        return
    if args.workload == "mixed":
        n_frames, steps = 4, max(1, min(args.steps, 3))  # (a 256-frame step would take minutes on the CPU)
        fps, per_step = cpu_reference_mixed(n_frames, steps)
        sample = (f"{n_frames} of the 256 frames x {steps} steps (CenterNet decode + per-frame box_decode/nms/"
                  "assemble_mask/masked depth mean), torch-CPU oracle port")
        workload = WORKLOAD_MIXED
    else:
        n_frames, steps = B_PER_GPU, max(1, args.steps)
        fps, per_step = cpu_reference_sample(n_frames, steps, warmup=max(1, min(args.warmup, 3)))
        sample = (f"{n_frames} frames x {steps} steps of decode+encode (the full batch), torch-CPU oracle port; the "
                  "decode leg is the tensor part only (no per-detection Python loop), which flatters the CPU")
        workload = WORKLOAD
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.workload == "mixed" else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "frames_per_step": n_frames},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------------------------------------
def median(xs):
    s = sorted(xs)
    return s[len(s) // 2]


def time_kernel(fn, reps=7, warmup=3, inner=4):
    """Median device time (us) per call of fn() over `reps` timings of `inner` back-to-back calls, CUDA events on the
    current stream, behind a short device-side spin (one call between two events on an idle stream would time the host's
    launch path, not the device).
    The YOLACT inputs are several times larger than the 126 MB L2, so every launch streams from HBM."""
    for _ in range(warmup):
        r = fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        torch.cuda._sleep(1_000_000)  # ~0.5 ms of device-side spin: the host enqueues all `inner` calls meanwhile
        e0.record()
        for _ in range(inner):
            r = fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / inner)
    return median(ts), r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--blocks", type=int, default=5, help="the K-step block is timed this many times; the median is reported")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="centernet", choices=["centernet", "mixed"])
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-yolact", action="store_true", help="skip the configs[2] kernel timings")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    args.blocks = max(5, args.blocks)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    numa_node, n_cores = pin_to_gpu_numa_node(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist_mod.init_process_group("nccl", device_id=device)
        dist = dist_mod

    import tauv_vision_b200 as tv
    lib = tv.load_library()
    assert lib.tauv_check_device() == 0, lib.tauv_last_error()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = SimpleNamespace(args=args, rank=rank, local_rank=local_rank, world=world, device=device, dist=dist,
                          barrier=barrier, numa_node=numa_node, n_cores=n_cores)
    if args.workload == "mixed":
        run_mixed(ctx)
    else:
        run_centernet(ctx)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------
# configs[1] (+ configs[2] kernels)
# ---------------------------------------------------------------------------------------------------------------------
def run_centernet(ctx):
    from tauv_vision_b200.centernet.model import decode as D
    from tauv_vision_b200.centernet.model import loss as L
    args, device, dist, world, rank = ctx.args, ctx.device, ctx.dist, ctx.world, ctx.rank
    mc = SimpleNamespace(in_h=IN_HW, in_w=IN_HW, downsample_ratio=2 ** DOWNSAMPLES, out_h=H, out_w=W)
    tc = SimpleNamespace(keypoint_heatmap_sigma=SIGMA)
    oc = SimpleNamespace(n_labels=C)
    logits, size, offset, truth = make_inputs(device, 1234 + rank)
    pred = SimpleNamespace(heatmap=logits, size=size, offset=offset, depth=None)
    K, NB = args.steps, args.blocks

    # outputs are allocated once and overwritten every step (a real pipeline re-uses its buffers too); the truth
    # tensors are already in the dtypes / layout the encoder takes, so a step is two kernel launches and nothing else
    det_buf = D.decode_packed(pred, mc, K_DET, THR)
    tgt_buf = torch.empty((B_PER_GPU, C, H, W), dtype=torch.float32, device=device)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K * NB)]

    def step(events=None):
        if events is None:
            det = D.decode_packed(pred, mc, K_DET, THR, out=det_buf)
            L.generate_heatmap(truth, mc, tc, oc, out=tgt_buf)
        else:
            events[0].record()
            det = D.decode_packed(pred, mc, K_DET, THR, out=det_buf)
            events[1].record()
            L.generate_heatmap(truth, mc, tc, oc, out=tgt_buf)
            events[2].record()
        return det

    for _ in range(args.warmup):
        step()
    ctx.barrier()
    block_ms = []
    with ClockSampler(ctx.local_rank) as clocks:
        for blk in range(NB):
            ctx.barrier()
            start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            start.record()
            for i in range(K):
                det = step(ev[blk * K + i])
            end.record()
            ctx.barrier()
            block_ms.append(start.elapsed_time(end))
    t_dec = sum(e[0].elapsed_time(e[1]) for e in ev) / len(ev)
    t_enc = sum(e[1].elapsed_time(e[2]) for e in ev) / len(ev)
    n_det_mean = float(det.count.float().mean())

    # the decode back to back with itself (the launch before it only reads: no dirty L2 lines to write back)
    for _ in range(3):
        D.decode_packed(pred, mc, K_DET, THR, out=det_buf)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(K):
        D.decode_packed(pred, mc, K_DET, THR, out=det_buf)
    e1.record()
    torch.cuda.synchronize()
    t_dec_iso = e0.elapsed_time(e1) / K

    # the same call on SMOOTH maps (box-filtered noise, 9 x 9 twice: ~1 % of the cells are peaks and the hot regions span
    # many blocks — what a trained head emits), back to back: the data-dependent half of the decode, for the record
    g_s = torch.Generator(device=device)
    g_s.manual_seed(99 + rank)
    sm = torch.randn((B_PER_GPU, C, H + 16, W + 16), device=device, generator=g_s)
    for _ in range(2):
        sm = torch.nn.functional.avg_pool2d(sm, 9, 1)
    sm = ((sm - sm.mean()) / sm.std() * 1.5 - 2.2).contiguous()
    pred_s = SimpleNamespace(heatmap=sm, size=size, offset=offset, depth=None)
    for _ in range(3):
        D.decode_packed(pred_s, mc, K_DET, THR, out=det_buf)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        D.decode_packed(pred_s, mc, K_DET, THR, out=det_buf)
    e1.record()
    torch.cuda.synchronize()
    t_dec_smooth = e0.elapsed_time(e1) / 20
    del sm, pred_s

    # the dominant kernel alone (launch 1 of the decode's 2): the block maxima of the batch, back to back
    from tauv_vision_b200 import _lib
    lib_ = __import__("tauv_vision_b200").load_library()
    ws_bm = torch.empty(lib_.tauv_heatmap_topk_workspace_bytes(B_PER_GPU, C, H, W, K_DET), dtype=torch.uint8, device=device)

    def block_maxima():
        rc = lib_.tauv_centernet_block_maxima(_lib.fptr(logits), B_PER_GPU, C, H, W, K_DET, ws_bm.data_ptr(), ws_bm.numel(),
                                              _lib.stream_ptr(device))
        assert rc == 0, lib_.tauv_last_error()

    for _ in range(3):
        block_maxima()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(K):
        block_maxima()
    e1.record()
    torch.cuda.synchronize()
    t_bm = e0.elapsed_time(e1) / K
    del ws_bm

    # the heatmap term of the training loss fused with the target render: forward, forward + backward
    tc_f = SimpleNamespace(keypoint_heatmap_sigma=SIGMA, heatmap_focal_loss_a=2.0, heatmap_focal_loss_b=4.0)
    t_focal, _ = time_kernel(lambda: L.heatmap_focal_loss(logits, truth, mc, tc_f), reps=5, warmup=2)
    xg = logits.detach().clone().requires_grad_(True)

    def focal_fb():
        xg.grad = None
        L.heatmap_focal_loss(xg, truth, mc, tc_f).backward()

    t_focal_fb, _ = time_kernel(focal_fb, reps=5, warmup=2)
    del xg

    # ---- configs[2]: the YOLACT post-process kernels, same run, every rank ----
    yl = None
    if not args.no_yolact:
        yl = time_yolact(device, 4321 + rank)

    # ---- e2e: pinned host inputs in (one staging buffer), packed detections + target checksum out, every step ----
    e2e = e2e_centernet(ctx, logits, size, offset, truth, mc, tc, oc)

    # ---- max over ranks ----
    vals = [median(block_ms), e2e["ms"], t_dec, t_enc, t_dec_iso, t_bm] + block_ms
    if yl is not None:
        vals += [yl["detect_us"], yl["mask_us"], yl["mask_depth_us"], yl["scores_us"], yl["match_us"],
                 yl["mask_binary_nearest_us"], yl["mask_binary_bilinear_us"]]
    times = torch.tensor(vals, device=device, dtype=torch.float64)
    h2d_rate = torch.tensor([e2e["h2d_gbs"]], device=device, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(h2d_rate, op=dist.ReduceOp.MIN)
    vals = times.tolist()
    med_ms, e2e_ms, t_dec, t_enc, t_dec_iso, t_bm = vals[:6]
    block_ms = vals[6:6 + NB]
    if rank != 0:
        return
    hbm_gbs, tf_peak, peak_src = peaks()
    frames = B_PER_GPU * world
    value = frames * K / (med_ms * 1e-3)
    e2e_value = frames * args.e2e_steps / (e2e_ms * 1e-3)
    achieved = algorithmic_bytes_decode(B_PER_GPU) / (t_dec * 1e-3) / 1e9
    traffic = None
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists():
        tj = json.loads(tp.read_text())
        if "block_max_kernel_bytes_per_launch" in tj:
            traffic = tj["block_max_kernel_bytes_per_launch"] + tj.get("select_kernel_bytes_per_launch", 0)
    n_blk = C * ((H + 7) // 8) * (W // 4)
    bm_bytes = 4 * B_PER_GPU * C * H * W + 4 * B_PER_GPU * (n_blk + (n_blk + 31) // 32)
    kernels = {
        "decode_us": t_dec * 1e3, "decode_isolated_us": t_dec_iso * 1e3,
        "decode_smooth_maps_us": t_dec_smooth * 1e3,
        "decode_smooth_maps_note": "the decode back to back on box-filtered noise (9 x 9, twice): hot regions that span "
                                   "many blocks, as a trained head emits; rank 0's figure",
        "gaussian_encode_us": t_enc * 1e3,
        "gaussian_encode_gbs": algorithmic_bytes_encode(B_PER_GPU) / (t_enc * 1e-3) / 1e9,
        "gaussian_encode_frac": algorithmic_bytes_encode(B_PER_GPU) / (t_enc * 1e-3) / 1e9 / hbm_gbs,
        "heatmap_focal_loss_us": t_focal, "heatmap_focal_loss_fwd_bwd_us": t_focal_fb,
        "heatmap_focal_loss_note": "focal_loss(sigmoid(logits), generate_heatmap(truth)).sum() in one pass over the logits, "
                                   "no target written (compute-bound: one exp, one log, one IEEE divide per cell); "
                                   "backward = one more pass that writes the gradient; rank 0's figures",
    }
    launches = 3 * K * NB + 2 * (K + 3) + (K + 3) + 22 * 2 + 22 * 3  # decode = 2 launches, encode 1, focal 2 (+1 backward)
    if yl is not None:
        det_us, mask_us, md_us, sc_us, match_us, mbn_us, mbb_us = vals[6 + NB:6 + NB + 7]
        nk = yl["n_keep_total"]
        kernels.update({
            "yolact_config": f"BASELINE configs[2]: B={B_PER_GPU}, {YL_N} priors, {YL_C1} classes, top_k {YL_TOPK}, "
                             f"{YL_P} protos {YL_HP}x{YL_HP}, depth {CAM_H}x{CAM_W}; mean n_keep {nk / B_PER_GPU:.1f}",
            "yolact_scores_us": sc_us,
            "yolact_scores_frac": B_PER_GPU * 4 * YL_N * YL_C1 / (sc_us * 1e-6) / 1e9 / hbm_gbs,
            "yolact_detect_us": det_us,
            "yolact_detect_bytes": algorithmic_bytes_yolact_detect(B_PER_GPU, nk),
            "yolact_detect_frac": algorithmic_bytes_yolact_detect(B_PER_GPU, nk) / (det_us * 1e-6) / 1e9 / hbm_gbs,
            "mask_us": mask_us,
            "mask_bytes": algorithmic_bytes_mask(B_PER_GPU, nk),
            "mask_hbm_frac": algorithmic_bytes_mask(B_PER_GPU, nk) / (mask_us * 1e-6) / 1e9 / hbm_gbs,
            "mask_tflops_useful": 2.0 * nk * YL_P * YL_HP * YL_HP / (mask_us * 1e-6) / 1e12,
            "mask_tensor_pct": 100.0 * (2.0 * nk * YL_P * YL_HP * YL_HP / (mask_us * 1e-6) / 1e12) / tf_peak,
            "mask_depth_us": md_us,
            "mask_depth_bytes": algorithmic_bytes_mask_depth(B_PER_GPU, nk),
            "mask_depth_hbm_frac": algorithmic_bytes_mask_depth(B_PER_GPU, nk) / (md_us * 1e-6) / 1e9 / hbm_gbs,
            "yolact_frames_per_s_detect_plus_mask_depth": B_PER_GPU / ((det_us + md_us) * 1e-6),
            "mask_binary_nearest_us": mbn_us, "mask_binary_bilinear_us": mbb_us,
            "mask_binary_bytes": algorithmic_bytes_mask_binary(B_PER_GPU, nk),
            "mask_binary_nearest_frac": algorithmic_bytes_mask_binary(B_PER_GPU, nk) / (mbn_us * 1e-6) / 1e9 / hbm_gbs,
            "mask_binary_bilinear_frac": algorithmic_bytes_mask_binary(B_PER_GPU, nk) / (mbb_us * 1e-6) / 1e9 / hbm_gbs,
            "match_anchors_us": match_us,
            "match_anchors_bytes": 16 * YL_N + 16 * B_PER_GPU * 16 + B_PER_GPU * YL_N * 30,
            "match_anchors_frac": (16 * YL_N + 16 * B_PER_GPU * 16 + B_PER_GPU * YL_N * 30) / (match_us * 1e-6) / 1e9 / hbm_gbs,
        })
        kernels.update({
            "yolact_loss_forward_us": yl["loss_forward_us"], "yolact_loss_fwd_bwd_us": yl["loss_fwd_bwd_us"],
            "yolact_loss_note": f"loss(prediction, truth, config) of yolact/model/loss.py:8-125 through the Python API "
                                f"(match + class/box terms with hard-negative mining + mask term; 9 launches forward, 13 "
                                f"with backward) at configs[2]'s shapes, 16 truths per frame, {yl['loss_positives']} "
                                f"positives in the batch; rank 0's figures",
            "pack_heads_us": yl["pack_heads_us"],
            "pack_heads_bytes": 2 * 4 * B_PER_GPU * YL_N * (YL_C1 + 4 + YL_P),
            "pack_heads_frac": 2 * 4 * B_PER_GPU * YL_N * (YL_C1 + 4 + YL_P) / (yl["pack_heads_us"] * 1e-6) / 1e9 / hbm_gbs,
        })
        launches += yl["launches"]
    cpu = None
    if not args.no_cpu_baseline and world == 1:  # (the CPU leg is reported at N = 1 only)
        fps, per = cpu_reference_sample(B_PER_GPU, 5)
        cpu = {"value": fps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{B_PER_GPU} frames x 5 reps of decode+encode ({per:.2f} s per batch), torch-CPU oracle port "
                         "(tensor part of decode only: no per-detection Python loop, which flatters the CPU)"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": med_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_gpu": B_PER_GPU, "global_batch": frames,
                   "l2": "inputs larger than L2 (335.5 MB logits + 335.5 MB targets per step vs 126 MB L2)",
                   "timing": f"median of {NB} blocks of {K} steps (CUDA events, max over ranks)",
                   "mean_detections_per_frame": n_det_mean},
        "block_ms": block_ms,
        "roofline": {"bound": "hbm", "kernel": "block_max_kernel + select_kernel (the whole decode call: peaks, top-k, boxes; "
                                                "2 launches, the second programmatically dependent)",
                     "achieved": achieved, "peak": hbm_gbs, "peak_source": peak_src, "unit": "GB/s",
                     "frac": achieved / hbm_gbs, "traffic": traffic,
                     "algorithmic_bytes": algorithmic_bytes_decode(B_PER_GPU), "us_per_launch": t_dec * 1e3,
                     "us_per_launch_isolated": t_dec_iso * 1e3,
                     "frac_isolated": algorithmic_bytes_decode(B_PER_GPU) / (t_dec_iso * 1e-3) / 1e9 / hbm_gbs,
                     "dominant_kernel": {"name": "block_max_kernel", "us_per_launch": t_bm * 1e3,
                                         "algorithmic_bytes": bm_bytes, "achieved": bm_bytes / (t_bm * 1e-3) / 1e9,
                                         "frac": bm_bytes / (t_bm * 1e-3) / 1e9 / hbm_gbs,
                                         "how": "tauv_centernet_block_maxima alone, events, back to back"},
                     "step": {"algorithmic_bytes": algorithmic_bytes_decode(B_PER_GPU) + algorithmic_bytes_encode(B_PER_GPU),
                              "us": med_ms * 1e3 / K,
                              "frac": (algorithmic_bytes_decode(B_PER_GPU) + algorithmic_bytes_encode(B_PER_GPU))
                                      / (med_ms * 1e-3 / K) / 1e9 / hbm_gbs,
                              "how": "decode + encode bytes over the whole timed step (median block / steps): the "
                                     "write-back of the encode's dirty lines is paid inside the decode, so the step "
                                     "as a whole is the bandwidth figure that does not depend on which launch pays it"},
                     "note": "frac is the whole decode call (both launches) inside the timed steps, where it follows the "
                             "target encode and streams while that kernel's dirty L2 lines (up to 126 MB) are written "
                             "back; isolated = the call back to back with itself"},
        "kernels": kernels,
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                "steps": args.e2e_steps, "ms_per_step": e2e_ms / args.e2e_steps,
                "h2d_gbs_per_rank_min": float(h2d_rate[0]), "copies_per_step": "1 H2D + 2 D2H",
                "numa_node": ctx.numa_node, "host_cores_pinned": ctx.n_cores},
        "gpu_launches": launches,
        "clocks": clocks.summary(),
    }
    print(json.dumps(line), flush=True)


def time_yolact(device, seed, B=B_PER_GPU):
    from tauv_vision_b200.yolact.model import masks, nms
    y = make_yolact_inputs(device, seed, B)
    out = torch.empty((B, YL_TOPK, YL_HP, YL_HP), dtype=torch.float32, device=device)
    lib = __import__("tauv_vision_b200").load_library()
    ws_d = torch.empty(lib.tauv_yolact_mask_depth_workspace_bytes(B, YL_HP, YL_HP, YL_TOPK), dtype=torch.uint8, device=device)
    sc_us, _ = time_kernel(lambda: nms.max_foreground_confidence(y.cls))
    det_us, det = time_kernel(lambda: nms.detect(y.cls, y.enc, y.anchor, y.cfg, YL_TOPK, YL_IOU, YL_CONF))
    mask_us, _ = time_kernel(lambda: masks.assemble_mask_batched(y.proto, y.coeff, det, out=out), reps=5, warmup=2)
    md_us, _ = time_kernel(lambda: masks.masked_depth_mean_batched(y.proto, y.coeff, det, y.depth, workspace=ws_d), reps=5, warmup=2)
    nk = int(det.n_keep.sum().item())
    del out
    # binarised masks at the camera resolution (yolact_node.py:135 + :178): one byte per camera pixel and kept mask
    out_b = torch.empty((B, YL_TOPK, CAM_H, CAM_W), dtype=torch.uint8, device=device)
    ws_b = torch.empty(lib.tauv_yolact_mask_binary_workspace_bytes(B, YL_HP, YL_HP, YL_TOPK), dtype=torch.uint8, device=device)
    mb_us = {}
    for mode in ("nearest", "bilinear"):
        mb_us[mode], _ = time_kernel(lambda: masks.assemble_mask_binary_batched(
            y.proto, y.coeff, det, (CAM_H, CAM_W), mode=mode, out=out_b, workspace=ws_b), reps=5, warmup=2)
    del out_b, ws_b
    # anchor matching (training-side target encode, loss.py:16-22,62-66): 16 truths per frame.  Its outputs (37 MB) fit in
    # L2, so a large READ between the launches evicts them without leaving dirty lines behind (a memset would)
    from tauv_vision_b200.yolact.model import loss as yl_loss
    from tests import synth
    tb, tvd = synth.truth_boxes(B, 16, seed=1)
    tb, tvd = tb.to(device), tvd.to(device)
    ts = []
    for _ in range(7):
        y.proto.sum()
        torch.cuda._sleep(400_000)  # (device-side spin: the wrapper's five allocations and its launch are enqueued meanwhile)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        yl_loss.match_anchors(y.anchor, tb, tvd, y.cfg)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    # the training-side loss (yolact/model/loss.py:8-125) at the same shapes: 16 truths per frame, half of them on
    # (jittered) priors so that positives exist, 550 x 550 segmentation maps; and the heads' level outputs -> [B,N,C]
    from tauv_vision_b200.yolact.model import prediction_head as yl_heads
    g = torch.Generator(device=device)
    g.manual_seed(seed + 1)
    pick = torch.randint(0, YL_N, (B, 8), device=device, generator=g)
    tb[:, :8] = y.anchor[0][pick] * (1 + 0.05 * torch.randn((B, 8, 4), device=device, generator=g).clamp(-1, 1))
    tvd[:, :8] = True
    tcls = torch.randint(1, YL_C1, (B, 16), device=device, generator=g)
    seg = torch.randint(0, 16, (B, 55, 55), device=device, generator=g, dtype=torch.uint8)   # (the dataset's type)
    seg = seg.repeat_interleave(10, 1).repeat_interleave(10, 2).contiguous()
    img_valid = torch.ones((B, 550, 550), dtype=torch.bool, device=device)
    y.cfg.negative_example_ratio = 3
    grads = [t.detach().clone().requires_grad_() for t in (y.cls, y.enc, y.coeff, y.proto)]
    pred = (grads[0], grads[1], grads[2], y.anchor, grads[3])
    truth = (tvd, tcls, tb, seg, img_valid)
    with torch.no_grad():
        lf_us, _ = time_kernel(lambda: yl_loss.loss(pred, truth, y.cfg), reps=5, warmup=2, inner=2)

    def loss_fb():
        for t in grads:
            t.grad = None
        yl_loss.loss(pred, truth, y.cfg)[0].backward()
    lfb_us, _ = time_kernel(loss_fb, reps=5, warmup=2, inner=2)
    n_positive = int(yl_loss.match_anchors(y.anchor, tb, tvd, y.cfg).positive_match.sum().item())
    del grads, pred
    sizes = synth.fpn_sizes(550, 550)
    lv = {c: [torch.randn((B, 3 * c, h, w), device=device, generator=g) for h, w in sizes] for c in (YL_C1, 4, YL_P)}
    hcfg = SimpleNamespace(n_classes=YL_C1 - 1, n_prototype_masks=YL_P)
    with torch.no_grad():
        ph_us, _ = time_kernel(lambda: yl_heads.pack_heads(lv[YL_C1], lv[4], lv[YL_P], hcfg), reps=5, warmup=2, inner=2)
    del lv
    return {"loss_forward_us": lf_us, "loss_fwd_bwd_us": lfb_us, "loss_positives": n_positive, "pack_heads_us": ph_us,
            "scores_us": sc_us, "detect_us": det_us, "mask_us": mask_us, "mask_depth_us": md_us, "n_keep_total": nk,
            "match_us": median(ts[2:]), "mask_binary_nearest_us": mb_us["nearest"],
            "mask_binary_bilinear_us": mb_us["bilinear"],
            "launches": 31 * 1 + 31 * 2 + 22 * 1 + 22 * 3 + 7 + 2 * 22 * 2 + 12 * 9 + 12 * 13 + 1 + 12 * 3}


def e2e_centernet(ctx, logits, size, offset, truth, mc, tc, oc):
    """The step through the public API from pinned host memory: every step copies its inputs in as ONE staging buffer
    and its results out (the packed detections: one buffer; the target checksum: 4 bytes), double-buffered over a copy
    and a compute stream (the H2D of step i+1 overlaps step i).  Returns ms for args.e2e_steps steps on the device
    clock, the bytes per step and the host->device rate this rank reached."""
    from tauv_vision_b200.centernet.model import decode as D
    from tauv_vision_b200.centernet.model import loss as L
    args, device = ctx.args, ctx.device
    # one host staging buffer: [logits | size (NCHW) | offset (NCHW) | center | label | valid], 256-byte aligned parts
    parts = [("logits", logits), ("size", size.permute(0, 3, 1, 2).contiguous()),
             ("offset", offset.permute(0, 3, 1, 2).contiguous()), ("center", truth.center), ("label", truth.label),
             ("valid", truth.valid)]
    layout, off = [], 0
    for name, tns in parts:
        nb = tns.numel() * tns.element_size()
        layout.append((name, tns.dtype, tuple(tns.shape), off, nb))
        off += (nb + 255) // 256 * 256
    h_stage = torch.empty((off,), dtype=torch.uint8).pin_memory()
    for (name, dtype, shape, o, nb), (_, tns) in zip(layout, parts):
        h_stage[o:o + nb].view(dtype).view(shape).copy_(tns.cpu())
    h2d = sum(nb for *_, nb in layout)
    copy_stream, comp_stream = torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)
    sets = []
    for _ in range(2):
        d_stage = torch.empty((off,), dtype=torch.uint8, device=device)
        v = {name: d_stage[o:o + nb].view(dtype).view(shape) for name, dtype, shape, o, nb in layout}
        sets.append(SimpleNamespace(
            stage=d_stage, v=v, det=None, tgt=torch.empty((B_PER_GPU, C, H, W), dtype=torch.float32, device=device),
            h_det=None, h_chk=torch.empty((), dtype=torch.float32).pin_memory(),
            copied=torch.cuda.Event(enable_timing=True), copy_start=torch.cuda.Event(enable_timing=True),
            done=torch.cuda.Event()))
    d2h = 0

    def launch(d):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(d.done)   # the previous user of this buffer set has finished
            d.copy_start.record(copy_stream)
            d.stage.copy_(h_stage, non_blocking=True)
            d.copied.record(copy_stream)
        with torch.cuda.stream(comp_stream):
            comp_stream.wait_event(d.copied)
            p = SimpleNamespace(heatmap=d.v["logits"], size=d.v["size"].permute(0, 2, 3, 1),
                                offset=d.v["offset"].permute(0, 2, 3, 1), depth=None)
            tr = SimpleNamespace(valid=d.v["valid"], label=d.v["label"], center=d.v["center"])
            d.det = D.decode_packed(p, mc, K_DET, THR, out=d.det)
            L.generate_heatmap(tr, mc, tc, oc, out=d.tgt)
            if d.h_det is None:
                d.h_det = torch.empty(d.det._storage.shape, dtype=torch.uint8).pin_memory()
            d.h_det.copy_(d.det._storage, non_blocking=True)     # D2H of the packed detections (one buffer)
            d.h_chk.copy_(d.tgt.sum(), non_blocking=True)        # D2H of the encode's result scalar
            d.done.record(comp_stream)

    def read(d):
        nonlocal d2h
        d.done.synchronize()
        d2h = d.h_det.numel() + 4
        return int(d.h_det[-16:].sum()), float(d.h_chk)

    def run(n):
        for i in range(n):
            launch(sets[i & 1])
            if i >= 1:
                read(sets[(i - 1) & 1])
        return read(sets[(n - 1) & 1])

    run(4)
    ctx.barrier()
    es, ee = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    es.record(copy_stream)
    run(args.e2e_steps)
    ee.record(comp_stream)
    ctx.barrier()
    ms = es.elapsed_time(ee)  # device clock from before the first H2D to after the last D2H
    copy_ms = min(s.copy_start.elapsed_time(s.copied) for s in sets)
    return {"ms": ms, "h2d": h2d, "d2h": d2h, "h2d_gbs": h2d / (copy_ms * 1e-3) / 1e9}


# ---------------------------------------------------------------------------------------------------------------------
# configs[3]: mixed CenterNet + YOLACT, 256 frames sharded over the GPUs
# ---------------------------------------------------------------------------------------------------------------------
def run_mixed(ctx):
    from tauv_vision_b200 import shard
    from tauv_vision_b200.centernet.model import decode as D
    from tauv_vision_b200.yolact.model import masks, nms
    args, device, dist, world, rank = ctx.args, ctx.device, ctx.dist, ctx.world, ctx.rank
    lo, hi = shard.frame_range(rank, world, MIXED_FRAMES)
    nf = hi - lo
    mc = SimpleNamespace(in_h=IN_HW, in_w=IN_HW, downsample_ratio=2 ** DOWNSAMPLES, out_h=H, out_w=W)
    logits, size, offset, _ = make_inputs(device, 1000 + lo, B=nf)
    pred = SimpleNamespace(heatmap=logits, size=size, offset=offset, depth=None)
    y = make_yolact_inputs(device, 2000 + lo, nf)
    lib = __import__("tauv_vision_b200").load_library()
    ws_d = torch.empty(lib.tauv_yolact_mask_depth_workspace_bytes(nf, YL_HP, YL_HP, YL_TOPK), dtype=torch.uint8, device=device)
    det_buf = D.decode_packed(pred, mc, K_DET, THR)
    K, NB = args.steps, args.blocks

    def step():
        cn = D.decode_packed(pred, mc, K_DET, THR, out=det_buf)
        yd = nms.detect(y.cls, y.enc, y.anchor, y.cfg, YL_TOPK, YL_IOU, YL_CONF)
        mean, cnt = masks.masked_depth_mean_batched(y.proto, y.coeff, yd, y.depth, workspace=ws_d)
        return cn, yd, mean, cnt

    for _ in range(args.warmup):
        out = step()
    ctx.barrier()
    block_ms = []
    with ClockSampler(ctx.local_rank) as clocks:
        for _ in range(NB):
            ctx.barrier()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(K):
                out = step()
            e.record()
            ctx.barrier()
            block_ms.append(s.elapsed_time(e))

    # e2e: the same step + device->host copies of the packed results + the host gather on rank 0, in frame order
    def e2e_step():
        cn, yd, mean, cnt = step()
        # the packed results stay on the device; ONE gather (NVLink) and ONE device->host copy on rank 0, frame order
        local = {"index": cn.index, "label": cn.label, "score": cn.score, "yx": cn.yx, "hw": cn.hw, "count": cn.count,
                 "yl_keep": yd.keep, "yl_n_keep": yd.n_keep, "yl_box": yd.box, "yl_score": yd.score,
                 "yl_class": yd.class_id, "yl_depth_mean": mean}
        nbytes = sum(v.numel() * v.element_size() for v in local.values())
        return shard.gather_frames(local, MIXED_FRAMES), nbytes

    e2e_step()
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        gathered, d2h = e2e_step()
    ctx.barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    if rank == 0:
        assert gathered["label"].shape[0] == MIXED_FRAMES and gathered["yl_keep"].shape[0] == MIXED_FRAMES, \
            "the host gather must return all 256 frames in frame order"

    vals = torch.tensor([median(block_ms), e2e_ms] + block_ms, device=device, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    vals = vals.tolist()
    if rank != 0:
        return
    med_ms, e2e_ms = vals[:2]
    nk = float(out[1].n_keep.float().mean())
    line = {
        "metric": METRIC, "value": MIXED_FRAMES * K / (med_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K,
        "warmup": args.warmup, "ms_per_step": med_ms / K, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_MIXED, "frames_total": MIXED_FRAMES, "frames_per_gpu": nf,
                   "l2": "inputs larger than L2 (per GPU at N = 8: 168 MB logits + 200 MB class logits + 312 MB prototypes)",
                   "timing": f"median of {NB} blocks of {K} steps (CUDA events, max over ranks)",
                   "mean_n_keep_rank0": nk},
        "block_ms": vals[2:],
        "cpu_baseline": None,
        "e2e": {"value": MIXED_FRAMES * args.e2e_steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": d2h, "steps": args.e2e_steps, "ms_per_step": e2e_ms / args.e2e_steps,
                "note": "inputs resident (the heads are produced on the owning GPU); every step gathers the packed "
                        "CenterNet + YOLACT results on rank 0 in frame order (shard.gather_frames: one NCCL gather of the "
                        "device buffers + one device->host copy) inside the timed region, wall clock between barriers"},
        "gpu_launches": (2 + 2 + 3) * K * NB,
        "clocks": clocks.summary(),
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
