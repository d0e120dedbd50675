#!/usr/bin/env python
"""bench.py — decoded frames/s of the detection-head hot path on N B200s, one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): CenterNet decode + Gaussian target encode of a batch of 64 synthetic
512x512 frames per GPU (stride 4 -> 80 classes x 128x128 heatmaps, top-100, 16 objects per frame).
One step = decode (1 kernel) + encode (1 kernel) over the batch.  Frames are independent, so N GPUs each
take their own 64 frames (weak scaling, no collective on the data path; NCCL only carries the timing scalar).

value     : frames/s with inputs resident in HBM, CUDA events around exactly K steps, max over ranks.
e2e       : the same step through the public Python API with pinned HOST inputs copied in and the packed
            detections + a target checksum copied out inside the timed region, double-buffered over two streams
            (the H2D copy of step i+1 overlaps step i's kernels and D2H); PCIe-bound: 352 MB in per step.
roofline  : the dominant kernel (tile_cluster_kernel: the whole decode, reads the 335.5 MB of logits once), timed
            live with events around its launch, against MEASURED_PEAKS.json.
cpu_baseline / --impl reference : the CPU oracle port (oracle/ref_port.py, torch-CPU with all host threads)
            on a bounded sample of the same workload.
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
METRIC, UNIT = "decoded_frames_per_sec", "frames/s"
WORKLOAD = "centernet_decode_topk100+gaussian_target_encode, batch 64 x [80,128,128] per GPU (BASELINE configs[1], stride 4)"


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def algorithmic_bytes_decode(B):
    # read the logits once + index/label/score (2*8+8+4) + size/offset gathers (16) + packed boxes (24)
    return 4 * B * C * H * W + B * K_DET * (28 + 16 + 24)


def algorithmic_bytes_encode(B):
    return 4 * B * C * H * W


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


def make_inputs(device, seed):
    """Synthetic head tensors generated on the device (N(-2.2,1.5) logits = the reference's heatmap bias init)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    logits = torch.randn((B_PER_GPU, C, H, W), device=device, generator=g) * 1.5 - 2.2
    size = (torch.rand((B_PER_GPU, 2, H, W), device=device, generator=g) * 0.3).permute(0, 2, 3, 1)
    offset = (torch.rand((B_PER_GPU, 2, H, W), device=device, generator=g) * 4).permute(0, 2, 3, 1)
    truth = SimpleNamespace(valid=torch.rand((B_PER_GPU, N_OBJ), device=device, generator=g) < 0.75,
                            label=torch.randint(0, C, (B_PER_GPU, N_OBJ), device=device, generator=g),
                            center=torch.rand((B_PER_GPU, N_OBJ, 2), device=device, generator=g))
    return logits, size, offset, truth


def cpu_reference_sample(n_frames: int, reps: int):
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

    step()  # warm-up
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    dt = time.perf_counter() - t0
    return n_frames * reps / dt, dt / reps


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; the reference is pure
    Python/torch and cannot travel to the GPU box) on the host cores.  Rank 0 only."""
    if rank != 0:
        return
    # bounded sample: 8 frames per step (~0.2 s on 8 cores; per-frame cost matches the 64-frame batch), fewer if
    # the requested number of steps would otherwise run for many minutes
    n_frames = 8 if args.steps <= 400 else 2
    fps, per_step = cpu_reference_sample(n_frames, max(1, args.steps))
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{n_frames} frames per step"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n_frames} frames x {max(1, args.steps)} steps of decode+encode, torch-CPU oracle port"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3

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
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist_mod.init_process_group("nccl", device_id=device)
        dist = dist_mod

    import tauv_vision_b200 as tv
    from tauv_vision_b200.centernet.model import decode as D
    from tauv_vision_b200.centernet.model import loss as L
    lib = tv.load_library()
    assert lib.tauv_check_device() == 0, lib.tauv_last_error()

    mc = SimpleNamespace(in_h=IN_HW, in_w=IN_HW, downsample_ratio=2 ** DOWNSAMPLES, out_h=H, out_w=W)
    tc = SimpleNamespace(keypoint_heatmap_sigma=SIGMA)
    oc = SimpleNamespace(n_labels=C)
    logits, size, offset, truth = make_inputs(device, 1234 + rank)
    pred = SimpleNamespace(heatmap=logits, size=size, offset=offset, depth=None)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    K = args.steps
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]

    # outputs are allocated once and overwritten every step (a real pipeline re-uses its buffers too); the truth
    # tensors are already in the dtypes / layout the encoder takes, so a step is three kernel launches and nothing else
    det_buf = D.decode_packed(pred, mc, K_DET, THR)
    tgt_buf = torch.empty((B_PER_GPU, C, H, W), dtype=torch.float32, device=device)

    def step(events=None):
        if events is None:
            det = D.decode_packed(pred, mc, K_DET, THR, out=det_buf)
            tgt = L.generate_heatmap(truth, mc, tc, oc, out=tgt_buf)
        else:
            events[0].record()
            det = D.decode_packed(pred, mc, K_DET, THR, out=det_buf)
            events[1].record()
            tgt = L.generate_heatmap(truth, mc, tc, oc, out=tgt_buf)
            events[2].record()
        return det, tgt

    for _ in range(args.warmup):
        step()
    barrier()
    with ClockSampler(local_rank) as clocks:
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for i in range(K):
            det, tgt = step(ev[i])
        end.record()
        barrier()
    total_ms = start.elapsed_time(end)
    t_dec = sum(e[0].elapsed_time(e[1]) for e in ev) / K
    t_enc = sum(e[1].elapsed_time(e[2]) for e in ev) / K
    n_keep_mean = float(det.count.float().mean())

    # ---- e2e: pinned host inputs in, packed detections + target checksum out, every step ----
    h_logits = logits.cpu().pin_memory()
    h_size = size.permute(0, 3, 1, 2).contiguous().cpu().pin_memory()      # NCHW storage as the model emits it
    h_offset = offset.permute(0, 3, 1, 2).contiguous().cpu().pin_memory()
    h_valid, h_label, h_center = (truth.valid.cpu().pin_memory(), truth.label.cpu().pin_memory(),
                                  truth.center.cpu().pin_memory())
    # Two device buffer sets and two streams: the H2D copy of step i+1 runs while step i computes and its results go
    # back (double buffering).  Every step still copies its own inputs in and its own results out, and the host reads
    # step i's results (detections + target checksum) from pinned memory before it launches step i+2.
    h2d = sum(t.numel() * t.element_size() for t in (h_logits, h_size, h_offset, h_valid, h_label, h_center))
    copy_stream, comp_stream = torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)
    sets = []
    for _ in range(2):
        d = SimpleNamespace(
            logits=torch.empty_like(logits), size=torch.empty_like(h_size, device=device),
            offset=torch.empty_like(h_offset, device=device),
            truth=SimpleNamespace(valid=torch.empty_like(truth.valid), label=torch.empty_like(truth.label),
                                  center=torch.empty_like(truth.center)),
            det=None, tgt=torch.empty((B_PER_GPU, C, H, W), dtype=torch.float32, device=device),
            host=None, h_chk=torch.empty((), dtype=torch.float32).pin_memory(),
            copied=torch.cuda.Event(), done=torch.cuda.Event())
        sets.append(d)
    d2h = 0

    def e2e_launch(d):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(d.done)   # the previous user of this buffer set has finished
            d.logits.copy_(h_logits, non_blocking=True)
            d.size.copy_(h_size, non_blocking=True)
            d.offset.copy_(h_offset, non_blocking=True)
            d.truth.valid.copy_(h_valid, non_blocking=True)
            d.truth.label.copy_(h_label, non_blocking=True)
            d.truth.center.copy_(h_center, non_blocking=True)
            d.copied.record(copy_stream)
        with torch.cuda.stream(comp_stream):
            comp_stream.wait_event(d.copied)
            p = SimpleNamespace(heatmap=d.logits, size=d.size.permute(0, 2, 3, 1), offset=d.offset.permute(0, 2, 3, 1),
                                depth=None)
            d.det = D.decode_packed(p, mc, K_DET, THR, out=d.det)
            L.generate_heatmap(d.truth, mc, tc, oc, out=d.tgt)
            if d.host is None:
                d.host = {k: torch.empty(getattr(d.det, k).shape, dtype=getattr(d.det, k).dtype).pin_memory()
                          for k in ("index", "label", "score", "yx", "hw", "count")}
            for k, hbuf in d.host.items():
                hbuf.copy_(getattr(d.det, k), non_blocking=True)      # D2H of the packed detections
            d.h_chk.copy_(d.tgt.sum(), non_blocking=True)            # D2H of the encode's result scalar
            d.done.record(comp_stream)

    def e2e_read(d):
        nonlocal d2h
        d.done.synchronize()
        d2h = sum(v.numel() * v.element_size() for v in d.host.values()) + 4
        return int(d.host["count"].sum()), float(d.h_chk)

    def e2e_run(n):
        for i in range(n):
            e2e_launch(sets[i & 1])
            if i >= 1:
                e2e_read(sets[(i - 1) & 1])
        return e2e_read(sets[(n - 1) & 1])

    e2e_run(4)
    barrier()
    es, ee = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    es.record(copy_stream)
    e2e_run(args.e2e_steps)
    ee.record(comp_stream)
    barrier()
    e2e_ms = es.elapsed_time(ee)  # device clock from before the first H2D to after the last D2H

    # ---- max over ranks ----
    times = torch.tensor([total_ms, e2e_ms, t_dec, t_enc], device=device, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, t_dec, t_enc = times.tolist()

    if rank == 0:
        hbm_gbs, peak_src = peaks()
        frames = B_PER_GPU * world
        value = frames * K / (total_ms * 1e-3)
        e2e_value = frames * args.e2e_steps / (e2e_ms * 1e-3)
        achieved = algorithmic_bytes_decode(B_PER_GPU) / (t_dec * 1e-3) / 1e9
        traffic = None
        tp = ROOT / "profiles" / "traffic.json"
        if tp.exists():
            traffic = json.loads(tp.read_text()).get("tile_cluster_kernel_bytes_per_launch")
        cpu = None
        if not args.no_cpu_baseline and world == 1:  # (the CPU leg is reported at N = 1 only)
            fps, per = cpu_reference_sample(B_PER_GPU, 5)
            cpu = {"value": fps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                   "sample": f"{B_PER_GPU} frames x 5 reps of decode+encode ({per:.2f} s per batch), torch-CPU oracle port"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_gpu": B_PER_GPU, "global_batch": frames,
                       "l2": "inputs larger than L2 (335.5 MB logits + 335.5 MB targets per step vs 126 MB L2)",
                       "mean_detections_per_frame": n_keep_mean},
            "roofline": {"bound": "hbm", "kernel": "tile_cluster_kernel<SIGMOID_PEAK> (whole decode: peaks, top-k, boxes)",
                         "achieved": achieved, "peak": hbm_gbs, "peak_source": peak_src, "unit": "GB/s",
                         "frac": achieved / hbm_gbs, "traffic": traffic,
                         "algorithmic_bytes": algorithmic_bytes_decode(B_PER_GPU), "us_per_launch": t_dec * 1e3},
            "kernels": {
                "decode_us": t_dec * 1e3, "gaussian_encode_us": t_enc * 1e3,
                "gaussian_encode_gbs": algorithmic_bytes_encode(B_PER_GPU) / (t_enc * 1e-3) / 1e9,
                "gaussian_encode_frac": algorithmic_bytes_encode(B_PER_GPU) / (t_enc * 1e-3) / 1e9 / hbm_gbs,
            },
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": args.e2e_steps, "ms_per_step": e2e_ms / args.e2e_steps},
            "gpu_launches": 2 * K,
            "clocks": clocks.summary(),
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
