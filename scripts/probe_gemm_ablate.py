"""Which resource binds the short-K GEMMs?  Needs the TOME_GEMM_ABLATE build:
   TOME_LIB_SUFFIX=_abl TOME_NVCC_EXTRA=-DTOME_GEMM_ABLATE python -m multi_modal_transformers_tokenmerge_b200.build
   TOME_LIB_SUFFIX=_abl python scripts/probe_gemm_ablate.py
Ablations give wrong results by design: 1 = no output stores, 2 = no operand loads, 4 = no MMAs, 8 = no epilogue at all,
16 = every store into the first 1024 rows (no HBM write stream), 32 = A operand from the first 1024 rows (no HBM read stream)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_modal_transformers_tokenmerge_b200 import ops, _lib
from bench_kernels import timeit
L = _lib.lib()
for (m, n, k, bmn) in [(133120, 1536, 384, 1), (137216, 1152, 384, 1), (133120, 384, 1536, 1), (137216, 2304, 768, 1)]:
    a = torch.randn(m, k, device="cuda").bfloat16()
    b = (torch.randn(k, n, device="cuda") if bmn else torch.randn(n, k, device="cuda")).bfloat16()
    out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    bias = torch.randn(n, device="cuda")
    for pair in (1, 0):
        L.tome_gemm_set_pair_mma(pair)
        row = []
        for abl in (0, 1, 2, 4, 8, 1 | 2, 2 | 8, 1 | 4, 4 | 8, 2 | 4, 1 | 2 | 4, 16, 32, 48):
            L.tome_gemm_set_ablate(abl)
            t = timeit(lambda: ops.gemm(a, b, m=m, n=n, k=k, b_major=bmn, out=out, bias=bias), iters=5)
            row.append(f"{abl}:{t*1e6:6.1f}")
        L.tome_gemm_set_ablate(0)
        print(f"{m}x{n}x{k} pair={pair}  us by ablation bits (1 no store, 2 no load, 4 no mma, 8 no epilogue):  " + "  ".join(row))
