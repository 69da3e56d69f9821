"""Host<->device copy rates of the box with pinned buffers of the bench's e2e sizes (18.4 MB up, 73.4 MB down)."""
import time, torch
dev = torch.device("cuda", 0)
up_h = torch.empty(512 * 512 * 35, dtype=torch.int16).pin_memory()
dn_h = [torch.empty(512 * 512 * 35, dtype=torch.float32).pin_memory() for _ in range(2)]
up_d = torch.empty_like(up_h, device=dev)
dn_d = [torch.empty(512 * 512 * 35, dtype=torch.float32, device=dev) for _ in range(2)]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(fn, n=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
def up():
    with torch.cuda.stream(s1):
        up_d.copy_(up_h, non_blocking=True)
def down():
    with torch.cuda.stream(s2):
        for h, d in zip(dn_h, dn_d):
            h.copy_(d, non_blocking=True)
def both():
    up(); down()
print(f"H2D 18.4 MB: {run(up):.2f} ms  ({18.35 / run(up):.1f} GB/s)")
print(f"D2H 73.4 MB: {run(down):.2f} ms  ({73.4 / run(down):.1f} GB/s)")
print(f"both directions at once: {run(both):.2f} ms")
t0 = time.perf_counter(); x = torch.empty(512 * 512 * 35, dtype=torch.float32).pin_memory(); print(f"pin_memory(36.7 MB): {(time.perf_counter() - t0) * 1e3:.1f} ms")
