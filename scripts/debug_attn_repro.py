"""Debug aid: repeatability of attention forward at full size, with the location and size of the differences."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_modal_transformers_tokenmerge_b200 import ops  # noqa: E402

torch.manual_seed(0)
B, T, H, D = int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 536, 6, 64
qkv = torch.randn(B, T, 3, H, D, device="cuda").bfloat16()
q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
ref = torch.nn.functional.scaled_dot_product_attention(q.transpose(1, 2).float(), k.transpose(1, 2).float(), v.transpose(1, 2).float()).transpose(1, 2)
outs = []
for _ in range(6):
    o, l = ops.attention_fwd(q, k, v)
    torch.cuda.synchronize()
    outs.append((o.clone(), l.clone()))
for i, (o, l) in enumerate(outs):
    err = (o.float() - ref).abs().amax(dim=(2, 3))          # [B, T]
    badrows = (err > 0.05).nonzero()
    print(f"run {i}: max err {err.max().item():.4f}, rows with err > 0.05: {len(badrows)}; first {badrows[:6].tolist()}")
    if i:
        dl = (l != outs[0][1])
        idx = dl.nonzero()
        print(f"   lse differs from run 0 at {len(idx)} entries; shape {tuple(l.shape)}; first {idx[:5].tolist()}")
        if len(idx):
            a, b = l[dl][:5].tolist(), outs[0][1][dl][:5].tolist()
            print("   values", a, b)
