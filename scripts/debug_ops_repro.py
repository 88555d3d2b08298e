"""Debug aid: run each hot op several times on identical full-size inputs and report bitwise differences."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_modal_transformers_tokenmerge_b200 import _lib as L, ops  # noqa: E402

torch.manual_seed(0)
B, T, C, H, D, F = 256, 536, 384, 6, 64, 1536
M = B * T
x = torch.randn(B, T, C, device="cuda").bfloat16()


def check(name, fn, n=4):
    outs = [fn() for _ in range(n)]
    torch.cuda.synchronize()
    outs = [o if isinstance(o, (tuple, list)) else (o,) for o in outs]
    bad = 0
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            if not torch.equal(a, b):
                bad += 1
                d = (a != b)
                rows = d.reshape(d.shape[0], -1).any(dim=1).nonzero().flatten()
                print(f"  {name}: differs in {int(d.sum())} elements, first rows {rows[:8].tolist()} of {d.shape[0]}")
    print(f"{name}: {'NON-DETERMINISTIC' if bad else 'ok'}")


wqkv = (torch.randn(C, 3 * C, device="cuda") * 0.05).bfloat16()
bq = torch.randn(3 * C, device="cuda") * 0.01
x2 = x.view(M, C)
check("gemm qkv bias", lambda: ops.gemm(x2, wqkv, m=M, n=3 * C, k=C, b_major=1, bias=bq))
w1 = (torch.randn(C, F, device="cuda") * 0.05).bfloat16()
b1 = torch.randn(F, device="cuda") * 0.01
check("gemm fc1 bias relu", lambda: ops.gemm(x2, w1, m=M, n=F, k=C, b_major=1, bias=b1, relu=True))
check("gemm fc1 bias relu drop", lambda: ops.gemm(x2, w1, m=M, n=F, k=C, b_major=1, bias=b1, relu=True, dropout_rate=0.1, dropout_seed=3, dropout_site=5))
h1 = ops.gemm(x2, w1, m=M, n=F, k=C, b_major=1, bias=b1, relu=True)
w2 = (torch.randn(F, C, device="cuda") * 0.05).bfloat16()
b2 = torch.randn(C, device="cuda") * 0.01
check("gemm fc2 bias resid", lambda: ops.gemm(h1, w2, m=M, n=C, k=F, b_major=1, bias=b2, residual=x2))
check("gemm out-proj bias resid drop", lambda: ops.gemm(x2, wqkv[:, :C].contiguous(), m=M, n=C, k=C, b_major=1, bias=b2, residual=x2,
                                                        dropout_rate=0.1, dropout_seed=3, dropout_site=6))
dy = torch.randn(M, C, device="cuda").bfloat16()
check("gemm dgrad gate", lambda: ops.gemm(dy, w2, m=M, n=F, k=C, gate=h1, gate_scale=1.0))
check("gemm dgrad plain", lambda: ops.gemm(h1, w1, m=M, n=C, k=F))
check("gemm wgrad", lambda: ops.gemm(x2, h1, m=C, n=F, k=M, a_major=1, b_major=1, out_dtype=torch.float32))
qkv = ops.gemm(x2, wqkv, m=M, n=3 * C, k=C, b_major=1, bias=bq).view(B, T, 3, H, D)
q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
check("attention fwd", lambda: ops.attention_fwd(q, k, v))
gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
check("layernorm fwd axis1", lambda: ops.layernorm_fwd(x, gamma, beta, axis=1))
kk = qkv[:, :, 1].reshape(B, T, H * D).contiguous()
check("sim_argmax", lambda: ops.sim_argmax(kk, heads=H, dim=D, batch_stride=T * H * D, token_stride=H * D, head_stride=D, tokens=T, batch=B)[:2])
nm, ni = ops.sim_argmax(kk, heads=H, dim=D, batch_stride=T * H * D, token_stride=H * D, head_stride=D, tokens=T, batch=B)[:2]
plan = ops.select_topr(nm, ni, T, 16)
check("select_topr", lambda: (lambda p: (p.edge_idx, p.dst_idx, p.row_map))(ops.select_topr(nm, ni, T, 16)))
size = torch.ones(B, T, device="cuda")
check("merge fwd", lambda: ops.merge_fwd(plan, x, size, 1)[:2])
x1, s1 = ops.merge_fwd(plan, x, size, 1)[:2]
dyo = torch.randn_like(x1)
check("merge bwd", lambda: ops.merge_bwd(plan, dyo, size, s1, 1))
y, mean, rstd = ops.layernorm_fwd(x, gamma, beta, axis=1)
dyl = torch.randn_like(x)
def ln_bwd():
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dx = ops.layernorm_bwd(x, dyl, gamma, mean, rstd, dg, db, axis=1)
    return dx, dg, db
check("layernorm bwd axis1", ln_bwd)
check("colsum", lambda: ops.colsum(h1))
check("dropout_colsum", lambda: ops.dropout_colsum(x2, 0.1, 3, 7))
o, l = ops.attention_fwd(q, k, v)
do = torch.randn(B, T, H, D, device="cuda").bfloat16()
check("attention bwd", lambda: ops.attention_bwd(q, k, v, o, l, do))
