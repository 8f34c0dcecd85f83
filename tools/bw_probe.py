import torch, time
x = torch.empty(6 * 1024**3 // 8, dtype=torch.float64, device='cuda')
for _ in range(3): x.fill_(1.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): x.fill_(2.0)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print('pure write (fill_) %.1f GB/s' % (x.numel() * 8 / ms / 1e6))
y = torch.empty_like(x)
e0.record()
for _ in range(10): y.copy_(x)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print('copy %.1f GB/s (read+write)' % (2 * x.numel() * 8 / ms / 1e6))
e0.record()
for _ in range(10): s = x.sum()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print('pure read (sum) %.1f GB/s' % (x.numel() * 8 / ms / 1e6))
