"""GPU-box probe: what the host link gives (pinned D2H / H2D by size, two streams, zero-copy stores).  Not a bench value."""
import subprocess, time, torch
print(subprocess.run("nvidia-smi -q -i 0 | grep -i -A12 'GPU Link Info'", shell=True, capture_output=True, text=True).stdout)
dev = torch.device("cuda:0")
for mb in (8, 33, 132, 512):
    n = mb << 20
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    for name, fn in (("D2H", lambda: h.copy_(d, non_blocking=True)), ("H2D", lambda: d.copy_(h, non_blocking=True))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10): fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 10
        print(f"{name} {mb} MiB: {dt * 1e3:.3f} ms  {n / dt / 1e9:.1f} GB/s", flush=True)
# two streams, halves
n = 33 << 20
d = torch.empty(n, dtype=torch.uint8, device=dev); h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def two():
    with torch.cuda.stream(s1): h[: n // 2].copy_(d[: n // 2], non_blocking=True)
    with torch.cuda.stream(s2): h[n // 2:].copy_(d[n // 2:], non_blocking=True)
for _ in range(3): two()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): two()
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
print(f"D2H 33 MiB on two streams: {dt * 1e3:.3f} ms {n / dt / 1e9:.1f} GB/s")
# D2H while H2D runs
def both():
    with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
d2 = torch.empty(n, dtype=torch.uint8, device=dev); h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
print(f"D2H + H2D 33 MiB concurrently: {dt * 1e3:.3f} ms")
