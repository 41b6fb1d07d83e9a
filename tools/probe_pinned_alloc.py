import time, torch
torch.cuda.init(); torch.zeros(1, device="cuda")
for mb in (64, 256, 1024, 4096):
    t0 = time.perf_counter(); b = torch.empty(mb << 20, dtype=torch.uint8, pin_memory=True); t1 = time.perf_counter()
    print("pin alloc", mb, "MB", round((t1 - t0) * 1e3, 1), "ms", round(mb / 1024 / (t1 - t0), 2), "GB/s", flush=True)
    del b
a = torch.empty(1 << 30, dtype=torch.uint8)
a.fill_(1)
d = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(a); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("pageable H2D 1 GB", round(1.0737 / (t1 - t0), 1), "GB/s")
    t0 = time.perf_counter(); a.copy_(d); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("pageable D2H 1 GB", round(1.0737 / (t1 - t0), 1), "GB/s")
b = torch.empty(1 << 30, dtype=torch.uint8)
for _ in range(2):
    t0 = time.perf_counter(); b.copy_(a); t1 = time.perf_counter()
    print("host memcpy 1 GB", round(1.0737 / (t1 - t0), 1), "GB/s", torch.get_num_threads(), "threads")
