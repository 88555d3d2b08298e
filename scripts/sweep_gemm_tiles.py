"""Every GEMM of one transformer layer (forward, data gradients, weight gradients, with the stack's epilogues) at the
octo-small / octo-base bench shapes, timed for each forced (CTA mode, tile width) and for the automatic choice, with a
numerical check of every forced combination against the automatic one.   python scripts/sweep_gemm_tiles.py [small|base]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_modal_transformers_tokenmerge_b200 import ops, _lib
from bench_kernels import timeit
L = _lib.lib()
which = sys.argv[1] if len(sys.argv) > 1 else "small"
C, F, r = (384, 1536, 16) if which == "small" else (768, 3072, 32)
B, T = 256, 536
M, Mo = B * T, B * (T - r)
dev = "cuda"
bf = lambda *s: torch.randn(*s, device=dev).bfloat16()

def cases():
    x, xo = bf(M, C), bf(Mo, C)
    wqkv, wo, w1, w2 = bf(C, 3 * C), bf(C, C), bf(C, F), bf(F, C)
    big, bigf = bf(M, 3 * C), bf(Mo, F)
    bits = torch.randint(-2**31, 2**31 - 1, (Mo, F // 32), device=dev, dtype=torch.int32)
    bias = lambda n: torch.randn(n, device=dev)
    yield "qkv fwd", dict(a=x, b=wqkv, m=M, n=3 * C, k=C, b_major=1, bias=bias(3 * C))
    yield "out fwd (+resid, drop)", dict(a=x, b=wo, m=M, n=C, k=C, b_major=1, bias=bias(C), residual=x, dropout_rate=0.1, dropout_seed=1, dropout_site=3)
    yield "mlp1 fwd (relu, drop, bits)", dict(a=xo, b=w1, m=Mo, n=F, k=C, b_major=1, bias=bias(F), relu=True, dropout_rate=0.1, dropout_seed=1, dropout_site=4, relu_bits_out=torch.empty_like(bits))
    yield "mlp2 fwd (+resid, drop)", dict(a=bigf, b=w2, m=Mo, n=C, k=F, b_major=1, bias=bias(C), residual=xo, dropout_rate=0.1, dropout_seed=1, dropout_site=5)
    yield "dm1 dgrad (gate bits)", dict(a=xo, b=w2, m=Mo, n=F, k=C, b_major=0, gate_bits=bits, gate_scale=1.1)
    yield "dh2 dgrad", dict(a=bigf, b=w1, m=Mo, n=C, k=F, b_major=0)
    yield "d attn_o dgrad", dict(a=x, b=wo, m=M, n=C, k=C, b_major=0)
    yield "dh dgrad", dict(a=big, b=wqkv, m=M, n=C, k=3 * C, b_major=0)
    yield "W2 wgrad", dict(a=bigf, b=xo, m=F, n=C, k=Mo, a_major=1, b_major=1, out_dtype=torch.float32)
    yield "W1 wgrad", dict(a=xo, b=bigf, m=C, n=F, k=Mo, a_major=1, b_major=1, out_dtype=torch.float32)
    yield "Wo wgrad", dict(a=x, b=x, m=C, n=C, k=M, a_major=1, b_major=1, out_dtype=torch.float32)
    yield "Wqkv wgrad", dict(a=x, b=big, m=C, n=3 * C, k=M, a_major=1, b_major=1, out_dtype=torch.float32)

tot_auto = tot_best = 0.0
for name, kw in cases():
    a, b = kw.pop("a"), kw.pop("b")
    odt = kw.pop("out_dtype", torch.bfloat16)
    out = torch.zeros(kw["m"], kw["n"], device=dev, dtype=odt)
    L.tome_gemm_force_tile(-1, 0)
    ref = ops.gemm(a, b, out=out.clone(), **kw).float()
    t_auto = timeit(lambda: ops.gemm(a, b, out=out, **kw), iters=5)
    row, best = [], (t_auto, "auto")
    for mode in (0, 1, 2, 3):
        for bn in ((128, 192, 256) if mode < 3 else (192, 256)):
            L.tome_gemm_force_tile(mode, bn)
            got = ops.gemm(a, b, out=out.clone().zero_(), **kw).float()
            err = (got - ref).abs().max().item() / (ref.abs().max().item() + 1e-9)
            t = timeit(lambda: ops.gemm(a, b, out=out, **kw), iters=5)
            row.append(f"m{mode}/bn{bn}:{t*1e6:6.1f}" + ("" if err < 2e-2 else f"(ERR {err:.1e})"))
            if t < best[0] and err < 2e-2:
                best = (t, f"m{mode}/bn{bn}")
    L.tome_gemm_force_tile(-1, 0)
    fl = 2.0 * kw["m"] * kw["n"] * kw["k"]
    tot_auto += t_auto; tot_best += best[0]
    print(f"{name:28s} {kw['m']}x{kw['n']}x{kw['k']}: auto {t_auto*1e6:6.1f} us ({fl/t_auto/1e12:6.0f} TF/s)  best {best[1]} {best[0]*1e6:6.1f}  | " + " ".join(row), flush=True)
print(f"layer total: auto {tot_auto*1e6:.0f} us, best-of-sweep {tot_best*1e6:.0f} us")
