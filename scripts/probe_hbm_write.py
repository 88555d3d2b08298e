"""HBM bandwidth by traffic mix on this GPU (CUDA events, buffers far larger than L2): write-only (fill), read-only (sum),
copy (1:1) and a 3:1 write:read mix like the K = 384 projection GEMMs.  Development probe; result kept in profiles/."""
import torch

def t(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(it):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) * 1e-3)
    return best

n = 1 << 30   # bf16 elements: 2 GiB
a = torch.empty(n, dtype=torch.bfloat16, device="cuda").normal_()
b = torch.empty(n, dtype=torch.bfloat16, device="cuda")
print(f"write-only  fill_      : {2*n/t(lambda: b.fill_(1.0))/1e9:8.1f} GB/s")
print(f"write-only  zero_      : {2*n/t(lambda: b.zero_())/1e9:8.1f} GB/s")
print(f"read-only   sum        : {2*n/t(lambda: a.view(torch.int16).sum())/1e9:8.1f} GB/s")
print(f"copy 1:1    copy_      : {4*n/t(lambda: b.copy_(a))/1e9:8.1f} GB/s (read + write bytes)")
# 1 read : 3 write -- out[3, n/4] = a[n/4] broadcast
a4 = a[: n // 4]
b3 = b[: 3 * (n // 4)].view(3, n // 4)
print(f"1 read : 3 write (expand copy): {2*(n//4)*4/t(lambda: b3.copy_(a4.expand(3, -1)))/1e9:8.1f} GB/s (read + write bytes)")
