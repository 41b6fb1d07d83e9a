import torch, time
dev = torch.device("cuda:0")
n = 512 << 20
h = torch.empty(n, dtype=torch.uint8, pin_memory=True); d = torch.empty(n, dtype=torch.uint8, device=dev)
h2 = torch.empty(n // 4, dtype=torch.uint8, pin_memory=True); d2 = torch.empty(n // 4, dtype=torch.uint8, device=dev)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
print("H2D GB/s", n / t(lambda: d.copy_(h, non_blocking=True)) / 1e9)
print("D2H GB/s", n / t(lambda: h.copy_(d, non_blocking=True)) / 1e9)
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
dt = t(both)
print("H2D with concurrent D2H(1/4 size): H2D GB/s", n / dt / 1e9)
import os; print("cpus", os.cpu_count())
