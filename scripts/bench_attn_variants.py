"""A/B timing of the attention kernels' variants (process-wide switches of the library) on the bench shapes.
Development tool:  python scripts/bench_attn_variants.py [fwd] [bwd]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_modal_transformers_tokenmerge_b200 import _lib, ops  # noqa: E402
from multi_modal_transformers_tokenmerge_b200.tokenizers.token_sequencer import sequence_groups  # noqa: E402
from bench_kernels import timeit  # noqa: E402

lib = _lib.lib()


def inputs(B, T, H):
    qkv = torch.randn(B, T, 3, H, 64, device="cuda").bfloat16()
    n_img = (T - 16) // 2 - 4
    g1, p1, allow, _ = sequence_groups(f"[TaskDescriptionPrefix{{16}}] [Image{{{n_img}}};Readout{{4}}]*2")
    rng = np.random.default_rng(0)
    kw = dict(size=torch.ones(B, T, device="cuda"), gid=torch.tensor(np.stack([rng.permutation(g1) for _ in range(B)])).cuda(),
              pos=torch.tensor(np.broadcast_to(p1, (B, len(g1))).copy()).cuda(), allow=torch.tensor(allow).cuda())
    return qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], kw


def main():
    which = sys.argv[1:] or ["fwd", "bwd"]
    shapes = [(256, 536, 6), (256, 440, 6), (256, 360, 6), (256, 536, 12), (32, 2080, 12), (16, 4096, 12)]
    for (B, T, H) in shapes:
        q, k, v, kw = inputs(B, T, H)
        fl = 4.0 * B * H * T * T * 64
        drop = dict(dropout_rate=0.1, dropout_seed=3, dropout_site=5)
        if "fwd" in which:
            ref = None
            for ts in (0, 1):
                lib.tome_attention_set_fwd_ts(ts)
                t0 = timeit(lambda: ops.attention_fwd(q, k, v, **kw), iters=5)
                t1 = timeit(lambda: ops.attention_fwd(q, k, v, **kw, **drop), iters=5)
                out, lse = ops.attention_fwd(q, k, v, **kw, **drop)
                if ref is None:
                    ref = (out.clone(), lse.clone())
                diff = (out.float() - ref[0].float()).abs().max().item()
                print(f"attn_fwd B{B} T{T} H{H} ts={ts}: masked {t0*1e6:8.1f} us {fl/t0/1e12:6.1f} TF/s | + dropout {t1*1e6:8.1f} us "
                      f"{fl/t1/1e12:6.1f} TF/s | max diff vs ts=0 {diff:.2e} lse equal {torch.equal(lse, ref[1])}", flush=True)
            lib.tome_attention_set_fwd_ts(1)
        if "bwd" in which:
            out, lse = ops.attention_fwd(q, k, v, **kw, **drop)
            do = torch.randn_like(out)
            ref = None
            for ts in (0, 1):
                lib.tome_attention_set_bwd_ts(ts)
                t0 = timeit(lambda: ops.attention_bwd(q, k, v, out, lse, do, **kw), iters=5)
                t1 = timeit(lambda: ops.attention_bwd(q, k, v, out, lse, do, **kw, **drop), iters=5)
                g = ops.attention_bwd(q, k, v, out, lse, do, **kw, **drop)
                if ref is None:
                    ref = [t.clone() for t in g]
                diff = max((a.float() - b.float()).abs().max().item() for a, b in zip(g, ref))
                print(f"attn_bwd B{B} T{T} H{H} ts={ts}: masked {t0*1e6:8.1f} us {2.5*fl/t0/1e12:6.1f} TF/s | + dropout {t1*1e6:8.1f} us "
                      f"{2.5*fl/t1/1e12:6.1f} TF/s | max diff vs ts=0 {diff:.2e}", flush=True)


if __name__ == "__main__":
    main()
