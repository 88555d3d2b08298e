"""Prints per-tensor relative errors of the native stack vs the oracle (development aid)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_gpu_stack as T
from multi_modal_transformers_tokenmerge_b200 import engine, ops
pkg = (ops, engine)
for (ln_axis, r, Lyr, b1s) in [(1, 4, 2, 8.0), (1, 6, 3, 8.0), (2, 4, 2, 8.0), (1, 0, 2, 8.0), (1, 4, 2, 0.0)]:
    eng, cfg, layers, pe, x, y, groups = T._build(pkg, 2, 2, 24, 128, 2, 256, Lyr, r, ln_axis)
    for d in layers: d["b1"] = d["b1"] + np.float32(b1s)
    eng.load_params(pe[0], layers)
    eng.zero_grad(); eng.forward(torch.tensor(x).cuda(), torch.tensor(y).cuda()); eng.backward(); torch.cuda.synchronize()
    no = []
    for l in range(Lyr):
        pl = eng.layer_plan(l); no.append(None if pl is None else (pl[0].cpu().numpy(), pl[1].cpu().numpy()))
    params, pet, xf, size, origin, loss, out, tr = T._oracle_run(eng, cfg, layers, pe, x, y, groups, no, torch.bfloat16)
    print(f"== ln_axis {ln_axis} r {r} L {Lyr} b1+{b1s}: final_x {T.rel_err(eng.final_x().float().cpu(), xf.detach()):.4f} loss {eng.loss[0].item():.5f} vs {loss.item():.5f}")
    g = eng.param_views(eng.grads.cpu())
    print("  pos_emb", f"{T.rel_err(g['pos_embedding'], pet.grad[0]):.4f}")
    for l in range(Lyr):
        p, gl = params[l], g["layers"][l]
        ref = dict(ln1_scale=p.ln1_scale.grad, ln1_bias=p.ln1_bias.grad, ln2_scale=p.ln2_scale.grad, ln2_bias=p.ln2_bias.grad,
                   wqkv=torch.cat([p.wq.grad, p.wk.grad, p.wv.grad], 1), bqkv=torch.cat([p.bq.grad, p.bk.grad, p.bv.grad]),
                   wo=p.wo.grad, bo=p.bo.grad, w1=p.w1.grad, b1=p.b1.grad, w2=p.w2.grad, b2=p.b2.grad)
        print("  L%d " % l + " ".join(f"{k}={T.rel_err(gl[k], v):.3f}" for k, v in ref.items()))
