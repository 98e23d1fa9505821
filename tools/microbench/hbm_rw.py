import torch, time
dev = torch.device("cuda", 0)
n = 1 << 30
x = torch.empty(n, dtype=torch.float32, device=dev)   # 4 GiB
y = torch.empty(n, dtype=torch.float32, device=dev)
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = t(lambda: x.zero_()); print("write-only  %.1f GB/s" % (4 * n / ms / 1e6))
ms = t(lambda: x.sum()); print("read-only   %.1f GB/s" % (4 * n / ms / 1e6))
ms = t(lambda: y.copy_(x)); print("copy (r+w)  %.1f GB/s" % (8 * n / ms / 1e6))
