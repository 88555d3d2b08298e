"""Where does the image front end differ from the oracle?  Per-token error map at the gato geometry, with the workspace
pre-filled with NaN bit patterns and with zeros (a read of memory the call did not write shows up as NaN / as a difference)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_modal_transformers_tokenmerge_b200 import _lib as L, model_configs, ops  # noqa: E402
from multi_modal_transformers_tokenmerge_b200.tokenizers.images import encode_patch_position  # noqa: E402
from oracle import tome_oracle as O  # noqa: E402
import ctypes as C  # noqa: E402

rng = np.random.default_rng(11)
tok = model_configs.build_image_tokenizer(model_configs.load("tokenizers/images/gato_resnet_octo"))
variables = tok.init(5, None)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 3
img = rng.integers(0, 256, size=(B, 2, 280, 280, 3)).astype(np.uint8)
p = O.image_tokenizer_params_from_flax(variables["params"], 2)
want = O.image_tokenizer_fwd(p, img.astype(np.float32), patch_size=56, position_interval=128, num_groups=32, normalize=True)
flat = tok.pack_params(variables["params"])
d = tok._desc(B, 2, torch.uint8, 1)
row, col = encode_patch_position(280, 56, 128, False)
rt, ct = torch.from_numpy(row).cuda(), torch.from_numpy(col).cuda()
nbytes = int(L.lib().tome_image_tokenizer_workspace_bytes(C.byref(d)))
imgd = torch.from_numpy(img).cuda()
outs = []
for fill in (0xFF, 0x00, 0xFF):
    ws = torch.full((nbytes + 256,), fill, dtype=torch.uint8, device="cuda")
    got = ops.image_tokenizer_fwd(imgd, flat, d, rt, ct, workspace=ws).float().cpu().numpy()
    outs.append(got)
    err = np.abs(got - want)
    tokerr = np.nan_to_num(err, nan=99.0).max(axis=-1).reshape(-1)
    bad = np.nonzero(tokerr > 0.1)[0]
    print(f"fill {fill:#x}: nan {int(np.isnan(got).sum())} max {np.nanmax(err):.4f} mean {np.nanmean(err):.5f}; bad tokens (flat index b*50+n*25+patch): {bad.tolist()[:40]}")
print("0xFF vs 0x00 identical:", np.array_equal(outs[0], outs[1], equal_nan=True), " 0xFF twice identical:", np.array_equal(outs[0], outs[2], equal_nan=True))
