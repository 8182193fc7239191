"""H2D / D2H bandwidth from pinned memory on this box: one copy stream vs several (bench.py's e2e is bound by it)."""
import time
import torch

n = 256 << 20
src = [torch.empty(n // 4, dtype=torch.uint8, pin_memory=True) for _ in range(4)]
dst = [torch.empty(n // 4, dtype=torch.uint8, device="cuda") for _ in range(4)]
back = torch.empty(n // 2, dtype=torch.uint8, pin_memory=True)
dback = torch.empty(n // 2, dtype=torch.uint8, device="cuda")
streams = [torch.cuda.Stream() for _ in range(4)]


def run(k, with_d2h=False, iters=10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        for i in range(4):
            with torch.cuda.stream(streams[i % k]):
                dst[i].copy_(src[i], non_blocking=True)
        if with_d2h:
            with torch.cuda.stream(streams[3]):
                back.copy_(dback, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / iters
    return n / dt / 1e9, dt * 1e3


for k in (1, 2, 4):
    for d2h in (False, True):
        gbs, ms = run(k, d2h)
        print("h2d 256 MiB on %d stream(s)%s: %.1f GB/s (%.2f ms)" % (k, " + 128 MiB d2h" if d2h else "", gbs, ms))
