"""ncu target: one training step of a short (default 2-layer) ToMe stack at the bench shape, bracketed by
cudaProfilerStart/Stop so `ncu --profile-from-start off` captures exactly that step.

    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/step python scripts/prof_step.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import ACTION_DIM, CONFIGS, MAX_ACTION, SEQ  # noqa: E402
from multi_modal_transformers_tokenmerge_b200.engine import StackConfig, ToMeStackEngine  # noqa: E402
from multi_modal_transformers_tokenmerge_b200.tokenizers.token_sequencer import sequence_groups  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "octo_small"
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 2
c = CONFIGS[name]
gid, pos, allow, ro = sequence_groups(SEQ)
T0, B, C = len(gid), c["batch"], c["channels"]
cfg = StackConfig(batch=B, tokens=T0, channels=C, heads=c["heads"], head_dim=c["head_dim"], mlp_dim=c["mlp_dim"],
                  layers=layers, r=c["r"], ln_axis=1, num_groups=allow.shape[0], n_readout=len(ro), dropout_rate=0.1,
                  dropout_seed=1, attn_dropout_rate=0.1, head="continuous", head_features=ACTION_DIM, max_action=MAX_ACTION)
eng = ToMeStackEngine(cfg, gid=gid, pos=pos, allow=allow, readout_idx=ro)
eng.init_params(seed=1)
x = torch.randn(B, T0, C, device="cuda").bfloat16()
y = torch.rand(B, ACTION_DIM, device="cuda") * 2 - 1   # target actions (bench.py --loss continuous)


def step():
    eng.zero_grad()
    eng.forward(x, y)
    eng.backward()
    eng.adamw_step(lr=1e-4)


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok loss", float(eng.loss[0]))
