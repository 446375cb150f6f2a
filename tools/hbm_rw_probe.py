import torch
x = torch.empty(17179869184 // 4, dtype=torch.float32, device="cuda")
y = torch.empty_like(x)
def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
ms = t(lambda: x.fill_(1.0)); print("fill_ 17.18 GB:", round(ms, 3), "ms", round(17.18 / ms, 2), "TB/s write-only")
ms = t(lambda: x.zero_()); print("zero_ (memset):", round(ms, 3), "ms", round(17.18 / ms, 2), "TB/s")
ms = t(lambda: y.copy_(x)); print("copy 17.18 GB:", round(ms, 3), "ms", round(2 * 17.18 / ms, 2), "TB/s read+write")
ms = t(lambda: x.sum()); print("sum (read-only):", round(ms, 3), "ms", round(17.18 / ms, 2), "TB/s")
