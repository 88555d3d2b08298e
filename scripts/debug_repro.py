"""Debug aid: run the full-size stack forward twice and report where the two runs differ."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_modal_transformers_tokenmerge_b200 import engine  # noqa: E402
from multi_modal_transformers_tokenmerge_b200.tokenizers.token_sequencer import sequence_groups  # noqa: E402

drop = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
adrop = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
masked = int(sys.argv[3]) if len(sys.argv) > 3 else 1
gid, pos, allow, ro = sequence_groups("[TaskDescriptionPrefix{16}] [Image{256};Readout{4}]*2")
B, T, C, H, Dff, Lyr, r, A = 256, 536, 384, 6, 1536, 12, 16, 8
cfg = engine.StackConfig(batch=B, tokens=T, channels=C, heads=H, head_dim=64, mlp_dim=Dff, layers=Lyr, r=r,
                         num_groups=allow.shape[0] if masked else 0, n_readout=len(ro), dropout_rate=drop, attn_dropout_rate=adrop,
                         dropout_seed=7, head="continuous", head_features=A, max_action=1.0)
eng = engine.ToMeStackEngine(cfg, gid=gid if masked else None, pos=pos if masked else None, allow=allow if masked else None,
                             readout_idx=ro)
eng.init_params(1)
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(B, T, C, device="cuda", generator=g).bfloat16()
act = torch.rand(B, A, device="cuda", generator=g) * 2 - 1
runs = []
for _ in range(3):
    eng.zero_grad()
    eng.forward(x, act)
    torch.cuda.synchronize()
    plans = [tuple(t.clone() for t in eng.layer_plan(l)) for l in range(Lyr)]
    runs.append((eng.loss.clone(), eng.final_x().clone(), plans))
for k in (1, 2):
    a, b = runs[0], runs[k]
    bad_rows = (a[1] != b[1]).flatten(1).any(dim=1).nonzero().flatten().tolist()
    print(f"run {k}: final_x rows differing: {len(bad_rows)} {bad_rows[:20]}")
    for l in range(Lyr):
        for name, ta, tb in zip(("node_max", "node_idx", "edge_idx", "dst_idx"), a[2][l], b[2][l]):
            if not torch.equal(ta, tb):
                rows = (ta != tb).flatten(1).any(dim=1).nonzero().flatten().tolist()
                print(f"  layer {l} {name}: rows {rows[:12]} ({len(rows)})")
                break
        else:
            continue
        break
