import torch, time
x = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
d = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, a, b in (("h2d", d, x), ("d2h", x, d)):
    for _ in range(2): a.copy_(b, non_blocking=True)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(10): a.copy_(b, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(name, round(2560 / 1024 / dt, 1), "GiB/s")
