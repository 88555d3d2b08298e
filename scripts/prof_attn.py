"""One attention forward + backward at the bench shape with the group mask (target for ncu)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_modal_transformers_tokenmerge_b200 import ops
from multi_modal_transformers_tokenmerge_b200.tokenizers.token_sequencer import sequence_groups
B, T, H = 256, 536, 6
gid, pos, allow, ro = sequence_groups("[TaskDescriptionPrefix{16}] [Image{256};Readout{4}]*2")
rng = np.random.default_rng(0)
g = torch.tensor(np.stack([rng.permutation(gid) for _ in range(B)])).cuda()
p = torch.tensor(np.broadcast_to(pos, (B, T)).copy()).cuda()
a = torch.tensor(allow).cuda()
qkv = torch.randn(B, T, 3, H, 64, device="cuda").bfloat16()
q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
size = torch.ones(B, T, device="cuda") * 2
for _ in range(3):
    out, lse = ops.attention_fwd(q, k, v, gid=g, pos=p, allow=a, size=size)
    do = torch.randn_like(out)
    ops.attention_bwd(q, k, v, out, lse, do, gid=g, pos=p, allow=a, size=size)
torch.cuda.synchronize()
print("ok")
