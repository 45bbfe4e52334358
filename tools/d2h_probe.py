"""device->host copy bandwidth of this box: one copy vs the same bytes split over several streams, pinned torch memory (run on the GPU box)."""
import json
import torch

dev = torch.device("cuda", 0)
out = []
for mb in (8, 33, 133, 512):
    n = mb << 20
    src = torch.empty(n, dtype=torch.uint8, device=dev)
    dst = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    for parts in (1, 2, 4):
        streams = [torch.cuda.Stream(dev) for _ in range(parts)]
        step = n // parts
        best = 1e9
        for rep in range(6):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i, s in enumerate(streams):
                s.wait_event(a)
                with torch.cuda.stream(s):
                    dst[i * step:(i + 1) * step].copy_(src[i * step:(i + 1) * step], non_blocking=True)
            for s in streams:
                torch.cuda.current_stream().wait_stream(s)
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        out.append({"MB": mb, "streams": parts, "ms": best, "GBps": n / best / 1e6})
        print(json.dumps(out[-1]), flush=True)
# host->device for comparison
n = 133 << 20
src = torch.empty(n, dtype=torch.uint8, pin_memory=True)
dst = torch.empty(n, dtype=torch.uint8, device=dev)
best = 1e9
for rep in range(6):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); dst.copy_(src, non_blocking=True); b.record(); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b))
print(json.dumps({"h2d_MB": 133, "ms": best, "GBps": n / best / 1e6}))
