"""Worker of tests/test_gpu_dp.py: one rank of a 2-GPU data-parallel step (launched by torch.distributed.run)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_modal_transformers_tokenmerge_b200.engine import StackConfig, ToMeStackEngine  # noqa: E402
from multi_modal_transformers_tokenmerge_b200.parallel import DataParallelTrainer  # noqa: E402
from multi_modal_transformers_tokenmerge_b200.tokenizers.token_sequencer import sequence_groups  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gid, pos, allow, ro = sequence_groups("[TaskDescriptionPrefix{8}] [Image{40};Readout{2}]*2")
    B, T, C, A = 8, len(gid), 256, 8
    cfg = StackConfig(batch=B, tokens=T, channels=C, heads=4, head_dim=64, mlp_dim=512, layers=3, r=6, ln_axis=1,
                      num_groups=allow.shape[0], n_readout=len(ro), head="continuous", head_features=A, max_action=1.0)
    eng = ToMeStackEngine(cfg, gid=gid, pos=pos, allow=allow, readout_idx=ro)
    eng.init_params(seed=1)                                    # identical weights on every rank
    g = torch.Generator(device="cuda").manual_seed(7)          # every rank draws ALL shards, uses its own
    xs = [torch.randn(B, T, C, device="cuda", generator=g).bfloat16() for _ in range(world)]
    ys = [torch.rand(B, A, device="cuda", generator=g) * 2 - 1 for _ in range(world)]
    overlap = os.environ.get("TOME_DP_OVERLAP", "layer")
    tr = DataParallelTrainer(eng, overlap=overlap)
    # --- "layer": per-layer events, side-stream all-reduce overlapped with backward; "none": one all-reduce after backward.
    # No optimiser step yet (lr = 0 keeps the weights)
    tr.train_step(xs[rank], ys[rank], lr=0.0)
    torch.cuda.synchronize()
    reduced = eng.grads.clone()
    # --- the same sum computed locally: every shard through this rank's engine, gradients accumulated without any exchange
    eng.zero_grad()
    for s in range(world):
        eng.forward(xs[s], ys[s])
        eng.backward()
    torch.cuda.synchronize()
    local_sum = eng.grads.clone()
    rel = ((reduced.double() - local_sum.double()).norm() / local_sum.double().norm()).item()
    views_r, views_l = eng.param_views(reduced), eng.param_views(local_sum)
    worst = 0.0
    for l in range(cfg.layers):
        for k in views_r["layers"][l]:
            a, b = views_r["layers"][l][k].double(), views_l["layers"][l][k].double()
            worst = max(worst, ((a - b).norm() / (b.norm() + 1e-30)).item())
    for k in ("kernel", "bias"):
        a, b = views_r["head"][k].double(), views_l["head"][k].double()
        worst = max(worst, ((a - b).norm() / (b.norm() + 1e-30)).item())
    worst = max(worst, ((views_r["pos_embedding"].double() - views_l["pos_embedding"].double()).norm()
                        / views_l["pos_embedding"].double().norm()).item())
    # --- two real steps: the replicas must stay bit-identical
    for _ in range(2):
        tr.train_step(xs[rank], ys[rank], lr=1e-3)
    torch.cuda.synchronize()
    chk = eng.params.view(torch.int32).to(torch.int64).sum().reshape(1)
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("DPRESULT " + json.dumps({"world": world, "overlap": overlap, "rel_err_all": rel, "worst_tensor_rel_err": worst,
                                        "params_in_sync": bool(lo.item() == hi.item()), "grad_norm": local_sum.norm().item()}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
