"""Probe: how much of the step is launch gaps?  Capture one training step in a CUDA graph and compare replay time with
stream launches (N = 1; the captured AdamW bakes its step count, so this is a timing probe, not a training loop)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import CONFIGS, SEQ
from multi_modal_transformers_tokenmerge_b200.engine import StackConfig, ToMeStackEngine
from multi_modal_transformers_tokenmerge_b200.tokenizers.token_sequencer import sequence_groups

c = CONFIGS["octo_small"]
gid, pos, allow, ro = sequence_groups(SEQ)
T0, B, C = len(gid), c["batch"], c["channels"]
cfg = StackConfig(batch=B, tokens=T0, channels=C, heads=c["heads"], head_dim=c["head_dim"], mlp_dim=c["mlp_dim"],
                  layers=c["layers"], r=c["r"], ln_axis=1, num_groups=allow.shape[0], n_readout=len(ro), dropout_rate=0.1, dropout_seed=1)
eng = ToMeStackEngine(cfg, gid=gid, pos=pos, allow=allow, readout_idx=ro)
eng.init_params(seed=1)
x = torch.randn(B, T0, C, device="cuda").bfloat16()
y = torch.randn(B, len(ro), C, device="cuda")

def step():
    eng.zero_grad(); eng.forward(x, y); eng.backward(); eng.adamw_step(lr=1e-4)

def timeit(fn, n=10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for _ in range(3): step()
print("stream launches: %.3f ms/step" % timeit(step))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2): step()
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g):
    step()
for _ in range(3): g.replay()
print("graph replay   : %.3f ms/step" % timeit(g.replay))
