"""Pure-write HBM bandwidth on this box (what a store-only kernel like trajectory mode can reach):
torch fill_ / zero_ / cudaMemset of 1.06 GB and a read+write copy for comparison."""
import torch
n = (1 << 20) * 252
a = torch.empty(n, dtype=torch.float32, device="cuda")
b = torch.empty(n, dtype=torch.float32, device="cuda")
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps
for name, fn, nbytes in (("fill_", lambda: a.fill_(1.5), 4 * n), ("zero_", lambda: a.zero_(), 4 * n),
                         ("copy_ (r+w)", lambda: b.copy_(a), 8 * n), ("sum (read)", lambda: a.sum(), 4 * n)):
    s = t(fn)
    print(f"{name:14s} {s*1e6:8.1f} us  {nbytes/s/1e9:8.1f} GB/s")
