import torch, time
dev = torch.device("cuda")
n, nv = 65536, 8
best = torch.randn(n, nv, device=dev); e = torch.rand(n, device=dev) * 100
def f():
    val, idx = torch.topk(e, 20, largest=False)
    return torch.cat([best[idx], val[:, None]], 1).contiguous()
for _ in range(5): f()
torch.cuda.synchronize()
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(100): f()
b.record(); torch.cuda.synchronize()
print("topk+gather+cat: %.1f us GPU per call" % (a.elapsed_time(b) * 10))
t0 = time.perf_counter()
for _ in range(100): f()
t1 = time.perf_counter(); torch.cuda.synchronize()
print("host issue time: %.1f us per call" % ((t1 - t0) * 1e4))
def g():
    m = torch.argmin(e)
    return m
a.record()
for _ in range(100): g()
b.record(); torch.cuda.synchronize()
print("argmin: %.1f us" % (a.elapsed_time(b) * 10))
