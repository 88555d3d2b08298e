"""ncu target: token-axis LayerNorm forward / backward at the bench shape."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_modal_transformers_tokenmerge_b200 import ops
B, T, C = 256, 536, 384
x = torch.randn(B, T, C, device="cuda").bfloat16()
g = torch.ones(C, device="cuda"); b = torch.zeros(C, device="cuda")
dres = torch.randn_like(x)
for _ in range(3):
    y, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-6, 1)
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    ops.layernorm_bwd(x, y, g, mean, rstd, dg, db, dres, 1)
torch.cuda.synchronize()
print("ok")
