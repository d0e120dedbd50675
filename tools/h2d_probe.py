import torch, time
dev = torch.device("cuda", 0)
n = 352 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device=dev)
def run(parts, reps=10):
    streams = [torch.cuda.Stream() for _ in range(parts)]
    step = (n // parts + 255) // 256 * 256
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(reps):
        evs = []
        for i, s in enumerate(streams):
            s.wait_event(e0) if r == 0 else None
            with torch.cuda.stream(s):
                lo, hi = i * step, min(n, (i + 1) * step)
                d[lo:hi].copy_(h[lo:hi], non_blocking=True)
                ev = torch.cuda.Event(); ev.record(s); evs.append(ev)
        for ev in evs: torch.cuda.current_stream().wait_event(ev)
    e1.record(); torch.cuda.synchronize()
    return n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
for parts in (1, 2, 4, 8):
    run(parts, 3)
    print(parts, "streams: %.1f GB/s" % run(parts))
