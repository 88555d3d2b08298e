"""ncu target: the matching kernels at the bench shape (packed K of the qkv buffer, 6 heads)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_modal_transformers_tokenmerge_b200 import ops
B, T, H, D = 256, 536, 6, 64
qkv = torch.randn(B, T, 3, H, D, device="cuda").bfloat16()
kw = dict(heads=H, dim=D, batch=B, tokens=T, batch_stride=T * 3 * H * D, token_stride=3 * H * D, head_stride=D, offset_elems=H * D)
for _ in range(3):
    nm, ni, _ = ops.sim_argmax(qkv, **kw)
    plan = ops.select_topr(nm, ni, T, 16)
torch.cuda.synchronize()
print("ok")
