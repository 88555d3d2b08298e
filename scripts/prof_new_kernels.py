"""ncu target for the kernels added at the end of round 2: one forward of the image front end (36 batch rows x 2 images, one
pass) and one importance + top-k prune at the bench shape, bracketed by cudaProfilerStart/Stop.

    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/new_kernels python scripts/prof_new_kernels.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_modal_transformers_tokenmerge_b200 import model_configs, ops  # noqa: E402
from multi_modal_transformers_tokenmerge_b200.tokenizers.token_sequencer import TokenSequence  # noqa: E402

tok = model_configs.build_image_tokenizer(model_configs.load("tokenizers/images/gato_resnet_octo"))
variables = tok.init(5, None)
img = torch.randint(0, 256, (36, 2, 280, 280, 3), dtype=torch.uint8, device="cuda")
B, T, H, D = 256, 536, 6, 64
ts = TokenSequence("[TaskDescriptionPrefix{16}] [Image{256};Readout{4}]*2", "[TaskDescriptionPrefix{0}] [Image{8};Readout{0}]*2")
gid, pos = (torch.as_tensor(a).cuda()[None].expand(B, -1).contiguous() for a in ts.group_ids())
allow = torch.as_tensor(ts.allow_table()).cuda()
qkv = torch.randn(B, T, 3, H, D, device="cuda").bfloat16()
q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
x = torch.randn(B, T, H * D, device="cuda").bfloat16()
sets = ts.prune_sets()
start, ss, sn, sk = 0, [], [], []
for n, c in sets:
    ss.append(start); sn.append(n); sk.append(n - c); start += n


def run():
    tok.apply(variables, img, train=False)
    _, lse = ops.attention_fwd(q, k, v, gid=gid, pos=pos, allow=allow)
    for mode in ("received", "row_mean"):
        imp = ops.attention_importance(q, k, lse, mode, gid=gid, pos=pos, allow=allow)
    ops.topk_prune(x, imp, ss, sn, sk)


run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
