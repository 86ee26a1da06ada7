import torch, time
x = torch.empty(1<<31, dtype=torch.float32, device="cuda")  # 8 GiB
y = torch.empty(1<<31, dtype=torch.float32, device="cuda")
for name, fn, bytes_ in (("memset (write only)", lambda: x.zero_(), 8<<30), ("fill (write only)", lambda: x.fill_(1.5), 8<<30), ("copy (read+write)", lambda: y.copy_(x), 16<<30), ("sum (read only)", lambda: x.sum(), 8<<30)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/10
    print(f"{name}: {ms:.3f} ms  {bytes_/ms/1e6:.0f} GB/s")
