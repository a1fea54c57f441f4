import torch
dev = torch.device("cuda:0")
for mb in (26, 104, 418, 548):
    x = torch.empty(mb * 1000 * 1000 // 4, dtype=torch.float32, device=dev)
    y = torch.empty_like(x)
    for _ in range(5): x.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): x.zero_()
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 20 * 1e3
    e0.record()
    for _ in range(20): x.fill_(1.5)
    e1.record(); torch.cuda.synchronize()
    t2 = e0.elapsed_time(e1) / 20 * 1e3
    e0.record()
    for _ in range(20): y.copy_(x)
    e1.record(); torch.cuda.synchronize()
    t3 = e0.elapsed_time(e1) / 20 * 1e3
    print(f"{mb} MB: memset {t:.1f} us = {mb / t * 1e3:.0f} GB/s | fill kernel {t2:.1f} us = {mb / t2 * 1e3:.0f} GB/s | copy {t3:.1f} us = {2 * mb / t3 * 1e3:.0f} GB/s (r+w)")
