import torch, time
n = 1 << 28
h = torch.empty(n, dtype=torch.float32).pin_memory()
h2 = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device="cuda")
d2 = torch.empty(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h2.copy_(d2, non_blocking=True))):
    fn(); torch.cuda.synchronize(); t = time.perf_counter(); fn(); torch.cuda.synchronize()
    print(name, "GB/s", n * 4 / (time.perf_counter() - t) / 1e9)
torch.cuda.synchronize(); t = time.perf_counter()
with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
print("both directions: GB/s each", n * 4 / (time.perf_counter() - t) / 1e9)
