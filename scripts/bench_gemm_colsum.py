"""dm1 data-gradient GEMM (gate bits) with and without the column-sum epilogue, against the separate column-sum pass."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_modal_transformers_tokenmerge_b200 import ops
from bench_kernels import timeit
for (m, n, k) in [(133120, 1536, 384), (129024, 3072, 768), (137216, 384, 384)]:
    a = torch.randn(m, k, device="cuda").bfloat16()
    w = torch.randn(n, k, device="cuda").bfloat16()
    bits = torch.randint(-2**31, 2**31 - 1, (m, (n + 31) // 32), device="cuda", dtype=torch.int32)
    out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    part = torch.empty((m + 127) // 128, n, device="cuda")
    acc = torch.zeros(n, device="cuda")
    t0 = timeit(lambda: ops.gemm(a, w, m=m, n=n, k=k, out=out, gate_bits=bits, gate_scale=1.1), iters=5)
    t1 = timeit(lambda: ops.gemm(a, w, m=m, n=n, k=k, out=out, gate_bits=bits, gate_scale=1.1, colsum_partial=part), iters=5)
    t2 = timeit(lambda: ops.reduce_rows(part, acc, True), iters=5)
    t3 = timeit(lambda: ops.colsum(out, acc, True), iters=5)
    print(f"{m}x{n}x{k}: gemm {t0*1e6:6.1f} us, gemm + column sums {t1*1e6:6.1f} us, reduce_rows {t2*1e6:5.1f} us | separate colsum pass {t3*1e6:6.1f} us")
