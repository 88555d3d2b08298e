"""Generate tests/golden/*.npz by EXECUTING the reference's own source files (imported from /root/reference,
never copied) under the numpy shim of jax/flax in oracle/jax_shim/.   TEST INFRASTRUCTURE ONLY.

Run in the build container (the GPU box has no /root/reference):   python oracle/gen_golden.py
The committed .npz fixtures are what tests/ read; this script documents how they were made.

Reference entry points executed:
  multi_modal_transformers/tokenizers/token_compression.py:54-129  bipartite_soft_matching, merge, merge_wavg
  multi_modal_transformers/tokenizers/token_compression.py:15-46   compute_top_k_tokens
  multi_modal_transformers/tokenizers/token_sequencer.py:186-334   TokenSequence.generate_attention_mask,
                                                                   get_modality_idx
  multi_modal_transformers/action_heads/continuous.py:12-25        ContinuousActionHead.__call__
  multi_modal_transformers/action_heads/categorical.py:12-40       assign_bins, CategoricalActionHead.__call__
  multi_modal_transformers/action_heads/diffusion.py:16-64         cosine_beta_schedule, FourierFeatures, OctoDenoise
  multi_modal_transformers/attention_blocks/attention.py:20-39     MLPBlock (instantiated by OctoDenoise)
  multi_modal_transformers/attention_blocks/attention.py:41-119    Encoder1DBlock.__call__, AddPositionEmbedding,
                                                                   StackedEncoder1DBlock.__call__ (nn.scan over the block) with
                                                                   the config nodes of model_configs/attention_blocks/
                                                                   vanilla_decoder.yaml: the reference's OWN control flow --
                                                                   which LayerNorm instance feeds what, where the residuals and
                                                                   the (deterministic) dropouts sit, how the scan threads the
                                                                   carry and slices the stacked parameters -- with the Flax
                                                                   leaf modules (LayerNorm, SelfAttention, Dense) restated in
                                                                   oracle/jax_shim/flax/linen.py  -> encoder_blocks.npz
  multi_modal_transformers/tokenizers/images/image_tokenizer.py:35-309  image_to_patches, encode_patch_position (train=False),
                                                                   ResNetV2Block.__call__, ImageTokenizer.__call__ with config
                                                                   nodes of the form of model_configs/tokenizers/images/
                                                                   gato_resnet.yaml at small widths (the literal 28 224 x 768 Dense
                                                                   kernel alone would be 87 MB); Flax leaves (Conv, GroupNorm,
                                                                   max_pool, gelu, Embed, Dense) restated in the shim
                                                                   -> image_tokenizer.npz
  (the two loss expressions live inside the Octo class, which needs the whole model: octo.py:163-165 and :183-187 are
   restated here in numpy float64 on the executed heads' outputs)
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(HERE, "..", "tests", "golden")


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    sys.dont_write_bytecode = True
    sys.path.insert(0, os.path.join(HERE, "jax_shim"))
    sys.path.insert(0, REF)
    import jax.numpy as jnp  # the shim

    tc = _load("ref_token_compression", f"{REF}/multi_modal_transformers/tokenizers/token_compression.py")
    ts = _load("ref_token_sequencer", f"{REF}/multi_modal_transformers/tokenizers/token_sequencer.py")
    os.makedirs(OUT, exist_ok=True)

    # ---------------- matching + merge ----------------
    rng = np.random.default_rng(20261018)
    cases = [
        # name, B, T, Dm, C, r, class_token, distill_token, kind
        ("small", 2, 16, 8, 4, 3, False, False, "normal"),
        ("oddT", 3, 75, 16, 8, 10, False, False, "normal"),
        ("octo_lit", 2, 74, 256, 32, 16, False, False, "normal"),
        ("clamp", 2, 10, 4, 4, 99, False, False, "normal"),
        ("cls", 2, 33, 8, 4, 5, True, False, "normal"),
        ("distill", 2, 32, 8, 4, 5, False, True, "normal"),
        ("cls_distill", 2, 41, 8, 4, 7, True, True, "normal"),
        ("ties", 2, 24, 2, 4, 5, False, False, "ties"),
        ("all_to_one", 1, 32, 4, 4, 8, False, False, "all_to_one"),
        ("c2_layer0", 1, 536, 64, 16, 16, False, False, "normal"),
    ]
    out = {}
    names = []
    for name, B, T, Dm, C, r, cls, dis, kind in cases:
        if kind == "normal":
            metric = rng.standard_normal((B, T, Dm)).astype(np.float32)
        elif kind == "ties":  # few distinct directions -> many exactly tied scores
            dirs = np.array([[1, 0], [0, 1], [-1, 0], [3, 4]], np.float32)
            metric = dirs[rng.integers(0, 4, size=(B, T))]
        elif kind == "all_to_one":  # every even token's best match is odd token 3
            metric = rng.standard_normal((B, T, Dm)).astype(np.float32) * 0.01
            metric[:, ::2, :] += np.array([1, 0, 0, 0], np.float32)
            metric[:, 7, :] = np.array([5, 0, 0, 0], np.float32)
        x = rng.standard_normal((B, T, C)).astype(np.float32)
        merge = tc.bipartite_soft_matching(jnp.asarray(metric), r, cls, dis)
        x1, s1 = tc.merge_wavg(merge, jnp.asarray(x))
        xsum = merge(jnp.asarray(x), mode="sum")
        # second round on the merged output (sizes now non-trivial), re-using the first T1 metric rows
        T1 = x1.shape[1]
        metric2 = rng.standard_normal((B, T1, Dm)).astype(np.float32)
        merge2 = tc.bipartite_soft_matching(jnp.asarray(metric2), r, cls, dis)
        x2, s2 = tc.merge_wavg(merge2, x1, s1)
        # the closure's indices (token_compression.py:84-88) -- read from the closure cells of `merge`
        cells = dict(zip(merge.__code__.co_freevars, [c.cell_contents for c in merge.__closure__]))
        cells2 = dict(zip(merge2.__code__.co_freevars, [c.cell_contents for c in merge2.__closure__]))
        names.append(name)
        out[f"{name}/cfg"] = np.array([B, T, Dm, C, r, int(cls), int(dis)], np.int32)
        out[f"{name}/metric"] = metric
        out[f"{name}/x"] = x
        out[f"{name}/x1"] = np.asarray(x1)
        out[f"{name}/s1"] = np.asarray(s1)
        out[f"{name}/xsum"] = np.asarray(xsum)
        out[f"{name}/unm_idx"] = np.asarray(cells["unm_idx"])[..., 0].astype(np.int32)
        out[f"{name}/src_idx"] = np.asarray(cells["src_idx"])[..., 0].astype(np.int32)
        out[f"{name}/dst_idx"] = np.asarray(cells["dst_idx"])[..., 0].astype(np.int32)
        out[f"{name}/metric2"] = metric2
        out[f"{name}/x2"] = np.asarray(x2)
        out[f"{name}/s2"] = np.asarray(s2)
        out[f"{name}/src_idx2"] = np.asarray(cells2["src_idx"])[..., 0].astype(np.int32)
        out[f"{name}/dst_idx2"] = np.asarray(cells2["dst_idx"])[..., 0].astype(np.int32)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "token_compression.npz"), **out)
    print("token_compression.npz:", names)

    # ---------------- per-modality top-k pruning (token_compression.py:15-46) ----------------
    # written to its own file so the matching goldens above stay byte-identical
    prng = np.random.default_rng(20261019)
    pcases = [
        # name, T, C, tokenset_idx, tokenset_k, kind
        ("two_sets", 40, 8, ((0, 16), (16, 24)), (4, 6), "normal"),
        ("octo_like", 74, 16, ((0, 16), (16, 25), (45, 25)), (16, 10, 10), "normal"),   # prefix kept whole, images pruned
        ("gaps", 64, 4, ((8, 20), (40, 24)), (5, 24), "normal"),                          # sets need not tile the sequence; k == n
        ("ties", 32, 4, ((0, 32),), (7,), "ties"),                                        # equal scores: lower index first
        ("constant", 48, 4, ((0, 16), (16, 32)), (3, 5), "constant"),                     # the reference's own 1/T scores
        ("c2", 536, 32, ((0, 16), (16, 256), (272, 4), (276, 256), (532, 4)), (16, 192, 4, 192, 4), "normal"),
    ]
    po = {"names": np.array([c[0] for c in pcases])}
    for name, T, C, sets, ks, kind in pcases:
        emb = prng.standard_normal((T, C)).astype(np.float32)
        if kind == "normal":
            imp = prng.random(T).astype(np.float32)
        elif kind == "ties":
            imp = prng.integers(0, 4, size=T).astype(np.float32)
        else:
            imp = np.full(T, 1.0 / T, np.float32)
        kept = tc.compute_top_k_tokens(jnp.asarray(emb), jnp.asarray(imp), sets, ks)   # the reference function itself
        po[f"{name}/emb"], po[f"{name}/imp"] = emb, imp
        po[f"{name}/sets"], po[f"{name}/ks"] = np.asarray(sets, np.int32), np.asarray(ks, np.int32)
        po[f"{name}/kept"] = np.asarray(kept)
    np.savez_compressed(os.path.join(OUT, "token_pruning.npz"), **po)
    print("token_pruning.npz:", [c[0] for c in pcases])

    # ---------------- token sequence masks ----------------
    seqs = {
        "octo_base": "[TaskDescriptionPrefix{16}] [Image{25};Readout{4}]*2",
        "main_demo": "[TaskDescriptionPrefix{20}] [Image{10};Readout{10}]*2",
        "w3": "[TaskDescriptionPrefix{4}] [Image{6};Readout{2}]*3",
        "text": "[Text{5}] [Image{3};Readout{2}]*2",
        "c2": "[TaskDescriptionPrefix{16}] [Image{256};Readout{4}]*2",
    }
    mo = {"names": np.array(list(seqs))}
    for name, s in seqs.items():
        seq = ts.TokenSequence(s)
        mask = np.asarray(seq.generate_attention_mask(repeats=1, layer=0))[0]
        mo[f"{name}/seq"] = np.array(s)
        mo[f"{name}/mask"] = np.packbits(mask.astype(np.uint8), axis=-1)
        mo[f"{name}/T"] = np.array(mask.shape[0], np.int32)
        mo[f"{name}/readout_idx"] = np.asarray(seq.get_modality_idx("readouts")).astype(np.int32)
    np.savez_compressed(os.path.join(OUT, "token_sequencer.npz"), **mo)
    print("token_sequencer.npz:", list(seqs))

    # ---------------- action heads (continuous.py, categorical.py) ----------------
    cont = _load("ref_continuous", f"{REF}/multi_modal_transformers/action_heads/continuous.py")
    cat = _load("ref_categorical", f"{REF}/multi_modal_transformers/action_heads/categorical.py")
    hrng = np.random.default_rng(20261020)
    ho = {}
    ccases = [("cont_small", 3, 4, 16, 7, 1.0), ("cont_octo", 4, 8, 768, 8, 2.0), ("cont_saturated", 2, 4, 32, 5, 0.25)]
    for name, B, n, C, A, mx in ccases:
        readouts = hrng.standard_normal((B, n, C)).astype(np.float32)
        kernel = (hrng.standard_normal((C, A)) * np.sqrt(2.0 / C)).astype(np.float32)
        bias = (hrng.standard_normal(A) * 0.01).astype(np.float32)
        actions = (hrng.uniform(-mx, mx, size=(B, A))).astype(np.float32)
        dense = {"_target_": "flax.linen.Dense", "features": A, "kernel": kernel, "bias": bias}
        head = cont.ContinuousActionHead(max_action=mx, attention_pooling={}, dense=dense)
        pred = np.asarray(head(jnp.asarray(readouts)))                     # the reference module itself, [B, 1, A]
        p64 = np.squeeze(pred).astype(np.float64)                           # octo.py:163
        loss = np.sum(np.square(p64 - actions), axis=-1)                    # octo.py:165
        ho.update({f"{name}/readouts": readouts, f"{name}/kernel": kernel, f"{name}/bias": bias, f"{name}/actions": actions,
                   f"{name}/max_action": np.float32(mx), f"{name}/pred": pred, f"{name}/loss": loss.astype(np.float32)})
    kcases = [("cat_small", 3, 2, 3, 16, 8, 1.0), ("cat_octo", 4, 8, 1, 768, 256, 2.0), ("cat_edges", 2, 4, 2, 32, 4, 1.0)]
    for name, B, A, ts_, C, bins, mx in kcases:
        n = A * ts_
        readouts = hrng.standard_normal((B, n, C)).astype(np.float32)
        kernel = (hrng.standard_normal((C, bins)) * np.sqrt(2.0 / C)).astype(np.float32)
        bias = (hrng.standard_normal(bins) * 0.01).astype(np.float32)
        actions = (hrng.uniform(-mx, mx, size=(B, A))).astype(np.float32)
        if name == "cat_edges":   # values on bin edges, at both bounds and outside them
            actions = np.array([[-1.0, -0.5, 0.0, 0.5], [1.0, 1.5, -1.5, 0.999]], np.float32)
        dense = {"_target_": "flax.linen.Dense", "features": bins, "kernel": kernel, "bias": bias}
        head = cat.CategoricalActionHead(num_bins=bins, max_action=mx, action_space_dim=A, dense=dense)
        logits = np.asarray(head(jnp.asarray(readouts)))                    # the reference module itself, [B, A, bins]
        target_bin = np.asarray(cat.assign_bins(jnp.asarray(actions), (-mx, mx), bins))   # the reference function itself
        onehot = (target_bin[..., None] == np.arange(bins)).astype(np.float64)   # jax.nn.one_hot: out of range -> zeros
        z = logits.astype(np.float64)
        logsm = z - z.max(-1, keepdims=True) - np.log(np.exp(z - z.max(-1, keepdims=True)).sum(-1, keepdims=True))
        loss = -(onehot * logsm).sum(-1)                                    # optax.softmax_cross_entropy, octo.py:187
        ho.update({f"{name}/readouts": readouts, f"{name}/kernel": kernel, f"{name}/bias": bias, f"{name}/actions": actions,
                   f"{name}/max_action": np.float32(mx), f"{name}/cfg": np.array([A, bins], np.int32), f"{name}/logits": logits,
                   f"{name}/target_bin": target_bin.astype(np.int32), f"{name}/loss": loss.astype(np.float32)})
    # ---------------- diffusion head: schedule + denoiser network (diffusion.py:16-64) ----------------
    import flax.linen as nn
    import multi_modal_transformers.action_heads.diffusion as dif     # imported as a package module: its config nodes
    #                                                                   name FourierFeatures / MLPBlock by _target_ path
    for steps in (8, 32):
        betas = np.asarray(dif.cosine_beta_schedule(steps))
        alphas = 1 - betas
        ho[f"diff_schedule{steps}/betas"] = betas
        ho[f"diff_schedule{steps}/alpha_hats"] = np.array([np.prod(alphas[: i + 1]) for i in range(steps)], np.float32)  # :88-92
    dcases = [("diff_small", 4, 3, 8, 32, 32, 48, 40, 64, 8), ("diff_mid", 8, 4, 8, 384, 128, 192, 128, 256, 32)]
    for name, B, n, A, C, F, Ht, To, H, steps in dcases:
        r_ = lambda *sh: (hrng.standard_normal(sh) * 0.1).astype(np.float32)  # noqa: E731
        p = dict(fourier_kernel=r_(F // 2, 1), tw1=r_(F, Ht), tb1=r_(Ht), tw2=r_(Ht, To), tb2=r_(To),
                 w1=(r_(A + To + C, H) * 0.5), b1=r_(H), w2=r_(H, A), b2=r_(A))
        dense = lambda k, b: {"_target_": "flax.linen.Dense", "features": k.shape[1], "kernel": k, "bias": b}  # noqa: E731
        mlp = lambda k0, b0, k1, b1: {"_target_": "multi_modal_transformers.attention_blocks.attention.MLPBlock",  # noqa: E731
                                      "dense": dense(k0, b0), "activation": {"_partial_": True, "_target_": "flax.linen.relu"},
                                      "norm": {"_target_": "flax.linen.Dropout", "rate": 0.1}, "dense_out": dense(k1, b1)}
        nn.Module.shim_params = {"fourier_kernel": p["fourier_kernel"]}
        te = {"_target_": "multi_modal_transformers.action_heads.diffusion.FourierFeatures", "output_dim": F,
              "kernel_init": {"_target_": "flax.linen.initializers.he_normal"}, "mlp_block": mlp(p["tw1"], p["tb1"], p["tw2"], p["tb2"])}
        den = dif.OctoDenoise(num_blocks=1, time_encoder=te, mlp_block=mlp(p["w1"], p["b1"], p["w2"], p["b2"]))
        readouts = hrng.standard_normal((B, n, C)).astype(np.float32)
        actions = hrng.uniform(-1, 1, (B, A)).astype(np.float32)
        noise = hrng.standard_normal((B, A)).astype(np.float32)
        time = hrng.integers(0, steps, (B, 1)).astype(np.int32)
        ah = ho[f"diff_schedule{steps}/alpha_hats"][time]                           # diffusion.py:131
        noisy = (np.sqrt(ah) * actions + np.sqrt(1 - ah) * noise).astype(np.float32)   # :132-134 (restated)
        emb = np.asarray(jnp.mean(jnp.asarray(readouts), axis=-2))                  # :107
        pred = np.asarray(den(jnp.asarray(noisy), jnp.asarray(time), jnp.asarray(emb)))   # the reference modules themselves
        loss = np.mean(np.sum(0.5 * (pred.astype(np.float64) - noise) ** 2, axis=-1))    # :141-142 (restated)
        for k_, v_ in p.items():
            ho[f"{name}/p/{k_}"] = v_
        ho.update({f"{name}/readouts": readouts, f"{name}/actions": actions, f"{name}/noise": noise, f"{name}/time": time,
                   f"{name}/steps": np.int32(steps), f"{name}/noisy": noisy, f"{name}/pred": pred, f"{name}/loss": np.float32(loss)})
    ho["diffusion"] = np.array([c[0] for c in dcases])
    ho["continuous"] = np.array([c[0] for c in ccases])
    ho["categorical"] = np.array([c[0] for c in kcases])
    np.savez_compressed(os.path.join(OUT, "action_heads.npz"), **ho)
    print("action_heads.npz:", list(ho["continuous"]), list(ho["categorical"]))


def gen_blocks():
    """encoder_blocks.npz: StackedEncoder1DBlock / Encoder1DBlock of the reference, executed (attention.py:41-119)."""
    sys.dont_write_bytecode = True
    for pth in (os.path.join(HERE, "jax_shim"), REF):
        if pth not in sys.path:
            sys.path.insert(0, pth)
    import flax.linen as nn
    import jax.numpy as jnp
    import multi_modal_transformers.attention_blocks.attention as att
    ts = _load("ref_token_sequencer2", f"{REF}/multi_modal_transformers/tokenizers/token_sequencer.py")

    rng = np.random.default_rng(20261019)
    out = {}
    # (name, B, sequence grammar or None, T if no grammar, C, H, Dff, num_blocks, reduction axis)
    cases = [("lit_seq", 2, "[TaskDescriptionPrefix{4}] [Image{6};Readout{2}]*2", 0, 64, 2, 64, 2, 1),   # yaml as written: LN over tokens
             ("allones_1blk", 3, None, 17, 32, 4, 48, 1, 1),   # nn.merge_param refuses mask=None (attention.py:55): all-ones mask
             ("feature_ln", 2, "[TaskDescriptionPrefix{3}] [Image{5};Readout{1}]*3", 0, 64, 2, 128, 3, -1)]
    for name, B, seq, T, C, H, Dff, N, ax in cases:
        mask = None
        if seq is not None:
            tseq = ts.TokenSequence(seq)
            m = np.asarray(tseq.generate_attention_mask(repeats=1))          # [1, T, T] bool, the reference's own rule table
            T = m.shape[-1]
            mask = np.broadcast_to(m[None], (B, 1, T, T)).copy()             # octo.py:119 adds the batch axis
        else:
            mask = np.ones((B, 1, T, T), bool)
        D = C // H
        r_ = lambda *sh, s_=1.0: (rng.standard_normal(sh) * s_).astype(np.float32)  # noqa: E731
        blk = {"LayerNorm_0": {"scale": 1 + r_(N, C, s_=0.1), "bias": r_(N, C, s_=0.1)},
               "LayerNorm_1": {"scale": 1 + r_(N, C, s_=0.1), "bias": r_(N, C, s_=0.1)},
               "SelfAttention_0": {k_: {"kernel": r_(N, C, H, D, s_=(2.0 / C) ** 0.5), "bias": r_(N, H, D, s_=0.01)} for k_ in ("query", "key", "value")},
               "MLPBlock_0": {"Dense_0": {"kernel": r_(N, C, Dff, s_=(2.0 / C) ** 0.5), "bias": r_(N, Dff, s_=0.01)},
                              "Dense_1": {"kernel": r_(N, Dff, C, s_=(2.0 / Dff) ** 0.5), "bias": r_(N, C, s_=0.01)}}}
        blk["SelfAttention_0"]["out"] = {"kernel": r_(N, H, D, C, s_=(2.0 / C) ** 0.5), "bias": r_(N, C, s_=0.01)}
        tree = {"posembed_input": {"pos_embedding": r_(1, T, C, s_=0.02)}, "ScanEncoder1DBlock_0": blk}
        dense = lambda f: {"_target_": "flax.linen.Dense", "features": f, "use_bias": True,  # noqa: E731
                           "kernel_init": {"_target_": "flax.linen.initializers.he_normal"},
                           "bias_init": {"_target_": "flax.linen.initializers.normal"}}
        enc = {  # model_configs/attention_blocks/vanilla_decoder.yaml:4-59, widths scaled down
            "_target_": "multi_modal_transformers.attention_blocks.attention.Encoder1DBlock",
            "layer_norm": {"_target_": "flax.linen.LayerNorm", "epsilon": 1e-6, "reduction_axes": [ax], "feature_axes": [-1],
                           "dtype": None, "param_dtype": None},
            "dropout": {"_target_": "flax.linen.Dropout", "rate": 0.1},
            "self_attention": {"_target_": "flax.linen.SelfAttention", "num_heads": H, "qkv_features": C, "dropout_rate": 0.1,
                               "decode": False, "kernel_init": {"_target_": "flax.linen.initializers.he_normal"}, "use_bias": True,
                               "bias_init": {"_target_": "flax.linen.initializers.normal"}, "dtype": "float32", "param_dtype": "float32"},
            "mlp_block": {"_target_": "multi_modal_transformers.attention_blocks.attention.MLPBlock", "dense": dense(Dff),
                          "activation": {"_partial_": True, "_target_": "flax.linen.relu"},
                          "norm": {"_target_": "flax.linen.Dropout", "rate": 0.1}, "dense_out": dense(C)}}
        x = r_(B, T, C)
        stack = att.StackedEncoder1DBlock(num_blocks=N, encoder_1d_block=enc)
        with nn.shim_scope({"StackedEncoder1DBlock_0": tree}):
            y = np.asarray(stack(jnp.asarray(x), train=False, mask=None if mask is None else jnp.asarray(mask)))
        # one block alone (layer 0's parameters), the Encoder1DBlock entry point itself
        one = att.Encoder1DBlock(layer_norm=enc["layer_norm"], dropout=enc["dropout"], self_attention=enc["self_attention"],
                                 mlp_block=enc["mlp_block"])
        with nn.shim_scope({"Encoder1DBlock_0": nn._tree_index(blk, 0)}):
            y1, none = one(jnp.asarray(x), mask=None if mask is None else jnp.asarray(mask), train=False)
        assert none is None
        out[f"{name}/x"], out[f"{name}/y"], out[f"{name}/y_block0"] = x, y.astype(np.float32), np.asarray(y1, np.float32)
        out[f"{name}/meta"] = np.array([B, T, C, H, Dff, N, ax], np.int32)
        out[f"{name}/mask"] = mask
        if seq is not None:
            out[f"{name}/seq"] = np.array(seq)

        def flat(prefix, t):
            for k_, v_ in t.items():
                if isinstance(v_, dict):
                    flat(f"{prefix}/{k_}", v_)
                else:
                    out[f"{prefix}/{k_}"] = v_
        flat(f"{name}/params", tree)
    out["cases"] = np.array([c[0] for c in cases])
    np.savez_compressed(os.path.join(OUT, "encoder_blocks.npz"), **out)
    print("encoder_blocks.npz:", [c[0] for c in cases])


def gen_image_tokenizer():
    """image_tokenizer.npz: the reference's ImageTokenizer executed in evaluation mode (train=False: the training mode draws
    the position tokens from jax.random, which no stand-in can reproduce)."""
    sys.dont_write_bytecode = True
    for pth in (os.path.join(HERE, "jax_shim"), REF):
        if pth not in sys.path:
            sys.path.insert(0, pth)
    import flax.linen as nn
    it = _load("ref_image_tokenizer", f"{REF}/multi_modal_transformers/tokenizers/images/image_tokenizer.py")
    rng = np.random.default_rng(20261019)
    # name, B, N, H, patch, C_in, features, groups, embed, position_interval, num_blocks, normalize
    cases = [("two_frames", 2, 2, 56, 28, 3, 16, 4, 32, 16, 2, True),
             ("nine_patches_one_block", 1, 3, 84, 28, 3, 8, 2, 24, 128, 1, True),
             ("group_size_one_raw_pixels", 3, 1, 40, 20, 3, 32, 32, 16, 64, 2, False)]
    out = {}
    for (name, B, N, H, P, Cin, F, G, E, PI, NB, norm) in cases:
        o2 = (P - 12) // 2 + 1 - 2
        f32 = lambda a: a.astype(np.float32)  # noqa: E731
        conv = lambda kh, cin: {"kernel": f32(rng.standard_normal((kh, kh, cin, F)) * (0.05 if kh == 12 else 0.1)),  # noqa: E731
                                "bias": f32(rng.standard_normal(F) * 0.01)}
        ef = {"Conv_0": conv(12, Cin), "Dense_0": {"kernel": f32(rng.standard_normal((o2 * o2 * F, E)) * 0.05),
                                                    "bias": f32(rng.standard_normal(E) * 0.01)}}
        for i in range(NB):
            ef[f"GroupNorm_{i}"] = {"scale": f32(1 + 0.1 * rng.standard_normal(F)), "bias": f32(0.1 * rng.standard_normal(F))}
            ef[f"Conv_{i + 1}"] = conv(3, F)
        tree = {"embedding_function": ef,
                "image_row_position_embedding": {"embedding": f32(rng.standard_normal((PI, E)) * 0.1)},
                "image_col_position_embedding": {"embedding": f32(rng.standard_normal((PI, E)) * 0.1)}}
        node = lambda t, **kw: dict(_target_=t, **kw)  # noqa: E731
        cfg = dict(image_size=(H, H, Cin), patch_size=P, normalize=norm, position_interval=PI, rng_collection="patch_encoding",
                   embedding_dim=E,
                   row_position_embedding=node("flax.linen.Embed", name="image_row_position_embedding", num_embeddings=PI, features=E),
                   col_position_embedding=node("flax.linen.Embed", name="image_col_position_embedding", num_embeddings=PI, features=E),
                   resnet=node("ref_image_tokenizer.ResNetV2Block", num_blocks=NB,
                               input_conv=node("flax.linen.Conv", features=F, kernel_size=[12, 12], strides=[2, 2], padding="VALID", use_bias=True),
                               input_pool=dict(_partial_=True, _target_="flax.linen.max_pool", window_shape=[3, 3], strides=[1, 1], padding="VALID"),
                               resnet_norm=node("flax.linen.GroupNorm", num_groups=G, epsilon=1e-6),
                               resnet_activation=dict(_partial_=True, _target_="flax.linen.gelu"),
                               resnet_conv=node("flax.linen.Conv", features=F, kernel_size=[3, 3], strides=[1, 1], padding="SAME", use_bias=True),
                               output_dense=node("flax.linen.Dense", features=E)))
        img = rng.integers(0, 256, size=(B, N, H, H, Cin)).astype(np.uint8)
        tok = it.ImageTokenizer(**cfg)
        with nn.shim_scope({"ImageTokenizer_0": tree}):
            y = np.asarray(tok(img.astype(np.float32), train=False), np.float32)
        # the pieces on their own: patches of the first image, evaluation-mode position tokens
        out[f"{name}/patches00"] = np.asarray(it.image_to_patches(img[0, 0].astype(np.float32), P, norm), np.float32)
        rt, ct = it.encode_patch_position(img[0, 0].astype(np.float32), None, P, PI, False)
        out[f"{name}/row_tokens"], out[f"{name}/col_tokens"] = np.asarray(rt, np.int32), np.asarray(ct, np.int32)
        out[f"{name}/image"], out[f"{name}/out"] = img, y
        out[f"{name}/meta"] = np.array([B, N, H, P, Cin, F, G, E, PI, NB, int(norm)], np.int32)

        def flat(prefix, t):
            for k_, v_ in t.items():
                if isinstance(v_, dict):
                    flat(f"{prefix}/{k_}", v_)
                else:
                    out[f"{prefix}/{k_}"] = v_
        flat(f"{name}/params", tree)
    out["cases"] = np.array([c[0] for c in cases])
    np.savez_compressed(os.path.join(OUT, "image_tokenizer.npz"), **out)
    print("image_tokenizer.npz:", [c[0] for c in cases])


class _Cfg(dict):
    """omegaconf.DictConfig lets the reference write `self.query_map_input.kernel_init` (attention.py:141): attribute access."""
    __getattr__ = dict.__getitem__


def gen_attention_pooling():
    """attention_pooling.npz: the reference's MultiHeadAttentionPooling (attention.py:122-150) executed in evaluation mode with
    config nodes of the form of model_configs/action_heads/diffusion.yaml:6-51 (widths scaled down)."""
    sys.dont_write_bytecode = True
    for pth in (os.path.join(HERE, "jax_shim"), REF):
        if pth not in sys.path:
            sys.path.insert(0, pth)
    import flax.linen as nn
    import jax.numpy as jnp
    import multi_modal_transformers.attention_blocks.attention as att
    rng = np.random.default_rng(20261020)
    out = {}
    # name, B, readouts, E, heads, Dff, LayerNorm reduction axis (diffusion.yaml says [1]: the single pooled token)
    cases = [("yaml_axes", 3, 4, 64, 2, 96, 1), ("feature_ln", 2, 8, 128, 4, 128, -1), ("one_readout", 2, 1, 64, 1, 64, -1)]
    for name, B, n, E, H, Dff, ax in cases:
        D = E // H
        r_ = lambda *sh, s_=1.0: (rng.standard_normal(sh) * s_).astype(np.float32)  # noqa: E731
        tree = {"learnt_q_input": r_(1, 1, E, s_=0.5),
                "MultiHeadDotProductAttention_0": {k_: {"kernel": r_(E, H, D, s_=(2.0 / E) ** 0.5), "bias": r_(H, D, s_=0.01)} for k_ in ("query", "key", "value")},
                "LayerNorm_0": {"scale": 1 + r_(E, s_=0.1), "bias": r_(E, s_=0.1)},
                "MLPBlock_0": {"Dense_0": {"kernel": r_(E, Dff, s_=(2.0 / E) ** 0.5), "bias": r_(Dff, s_=0.01)},
                               "Dense_1": {"kernel": r_(Dff, E, s_=(2.0 / Dff) ** 0.5), "bias": r_(E, s_=0.01)}}}
        tree["MultiHeadDotProductAttention_0"]["out"] = {"kernel": r_(H, D, E, s_=(2.0 / E) ** 0.5), "bias": r_(E, s_=0.01)}
        he = {"_target_": "flax.linen.initializers.he_normal"}
        dense = lambda f: {"_target_": "flax.linen.Dense", "features": f, "use_bias": True, "kernel_init": he,  # noqa: E731
                           "bias_init": {"_target_": "flax.linen.initializers.normal"}}
        pool = att.MultiHeadAttentionPooling(
            query_map_input=_Cfg(kernel_init=he),
            dot_product_attention={"_target_": "flax.linen.MultiHeadDotProductAttention", "num_heads": H, "kernel_init": he},
            layer_norm={"_target_": "flax.linen.LayerNorm", "epsilon": 1e-6, "reduction_axes": [ax], "feature_axes": [-1]},
            mlp_block={"_target_": "multi_modal_transformers.attention_blocks.attention.MLPBlock", "dense": dense(Dff),
                       "activation": {"_partial_": True, "_target_": "flax.linen.relu"},
                       "norm": {"_target_": "flax.linen.Dropout", "rate": 0.1}, "dense_out": dense(E)})
        x = r_(B, n, E)
        with nn.shim_scope({"MultiHeadAttentionPooling_0": tree}):
            y = np.asarray(pool(jnp.asarray(x), train=False), np.float32)
        assert y.shape == (B, 1, E)
        out[f"{name}/x"], out[f"{name}/y"] = x, y
        out[f"{name}/meta"] = np.array([B, n, E, H, Dff, ax], np.int32)

        def flat(prefix, t):
            for k_, v_ in t.items():
                if isinstance(v_, dict):
                    flat(f"{prefix}/{k_}", v_)
                else:
                    out[f"{prefix}/{k_}"] = v_
        flat(f"{name}/params", tree)
    out["cases"] = np.array([c[0] for c in cases])
    np.savez_compressed(os.path.join(OUT, "attention_pooling.npz"), **out)
    print("attention_pooling.npz:", [c[0] for c in cases])


if __name__ == "__main__":
    main()
    gen_blocks()
    gen_image_tokenizer()
    gen_attention_pooling()
