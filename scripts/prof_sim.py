import sys, os, torch
sys.path.insert(0, "/root/repo")
from multi_modal_transformers_tokenmerge_b200 import ops
B, T, H = 256, 536, 6
qkv = torch.randn(B, T, 3, H, 64, device="cuda").bfloat16()
kwm = dict(heads=H, dim=64, batch=B, tokens=T, batch_stride=T * 3 * H * 64, token_stride=3 * H * 64, head_stride=64, offset_elems=H * 64)
for _ in range(5):
    ops.sim_argmax(qkv, **kwm)
torch.cuda.synchronize()
print("ok")
