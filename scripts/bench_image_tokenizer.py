"""Image front end (csrc/image_tokenizer.cu) at gato_resnet.yaml's geometry: time per forward with CUDA events and the
per-kernel breakdown (ncu launch list of the same command gives the kernel names).  python scripts/bench_image_tokenizer.py [B] [N]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_modal_transformers_tokenmerge_b200 import _lib as L  # noqa: E402
from multi_modal_transformers_tokenmerge_b200 import model_configs  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2
H, P, F, G, E, PI, NB = 280, 56, 64, 32, 768, 128, 2
tok = model_configs.build_image_tokenizer(model_configs.load("tokenizers/images/gato_resnet_octo"))
variables = tok.init(5, None)
img = torch.randint(0, 256, (B, N, H, H, 3), dtype=torch.uint8, device="cuda")
for _ in range(3):
    out = tok.apply(variables, img, train=False)
torch.cuda.synchronize()
lib = L.lib()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 10
e0.record()
for _ in range(steps):
    out = tok.apply(variables, img, train=False)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
L.check(lib.tome_profile_enable(4096))
out = tok.apply(variables, img, train=False)
torch.cuda.synchronize()
prof = L.profile_collect()
lib.tome_profile_disable()
o1, o2 = 23, 21
flops = 2.0 * B * N * 25 * (o1 * o1 * 432 * F + NB * o2 * o2 * 9 * F * F + o2 * o2 * F * E)
pix = B * N * H * H * 3
print(json.dumps({"batch": B, "images": N, "ms_per_forward": ms, "images_per_s": B * N / ms * 1e3, "tflops": flops / ms / 1e9,
                  "pixel_bytes": pix, "token_bytes": B * N * 25 * E * 2,
                  "by_tag": {k: {"ms": v[0], "work": v[1], "ops": v[2]} for k, v in prof.items() if v[2]}}))
