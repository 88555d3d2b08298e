"""Debug aid: V[k, d] = 1 iff d == k // 64, so out[q, j] is the probability mass row q puts on key tile j."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_modal_transformers_tokenmerge_b200 import ops  # noqa: E402

torch.manual_seed(0)
B, T, H, D = 256, 536, 6, 64
qkv = torch.randn(B, T, 3, H, D, device="cuda").bfloat16()
qkv[:, :, 2] = 0
for j in range((T + 63) // 64):
    qkv[:, j * 64:(j + 1) * 64, 2, :, j] = 1
q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
ref = torch.nn.functional.scaled_dot_product_attention(q.transpose(1, 2).float(), k.transpose(1, 2).float(), v.transpose(1, 2).float()).transpose(1, 2)
for it in range(4):
    o, l = ops.attention_fwd(q, k, v)
    torch.cuda.synchronize()
    err = (o.float() - ref)                                  # [B, T, H, D]
    bad = (err.abs().amax(dim=3) > 0.01).nonzero()
    print(f"run {it}: bad (b, t, h) rows: {len(bad)}")
    seen = set()
    for b_, t_, h_ in bad.tolist():
        key = (b_, h_, t_ // 32)
        if key in seen:
            continue
        seen.add(key)
        if len(seen) > 6:
            break
        print(f"  b={b_} h={h_} row={t_}: mass diff per tile {[round(x, 3) for x in err[b_, t_, h_, :9].tolist()]}  sum {o[b_, t_, h_, :9].float().sum().item():.4f}")
