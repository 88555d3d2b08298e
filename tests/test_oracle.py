"""Pins the CPU oracle (oracle/tome_oracle.py) against
  (1) golden vectors produced by executing the reference's own token_compression.py / token_sequencer.py
      under a numpy shim of jax (oracle/gen_golden.py -> tests/golden/*.npz), and
  (2) the hand-checked vector of SURVEY.md Appendix B,
plus the invariants of SURVEY §8c.  CPU only."""
import os

import numpy as np
import pytest

from oracle import tome_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TC = np.load(os.path.join(GOLD, "token_compression.npz"))
TS = np.load(os.path.join(GOLD, "token_sequencer.npz"))


@pytest.mark.parametrize("name", [str(n) for n in TC["names"]])
def test_matching_and_merge_against_reference_golden(name):
    B, T, Dm, C, r, cls, dis = [int(v) for v in TC[f"{name}/cfg"]]
    metric, x = TC[f"{name}/metric"], TC[f"{name}/x"]
    plan = O.bipartite_soft_matching(metric, r, bool(cls), bool(dis))
    assert plan.r == O.clamp_r(T, r, cls, dis)
    np.testing.assert_array_equal(plan.src_idx, TC[f"{name}/src_idx"])
    np.testing.assert_array_equal(plan.dst_idx, TC[f"{name}/dst_idx"])
    np.testing.assert_array_equal(plan.unm_idx, TC[f"{name}/unm_idx"])
    x1, s1 = O.merge_wavg(plan, x)
    np.testing.assert_array_equal(s1, TC[f"{name}/s1"])
    np.testing.assert_array_equal(x1, TC[f"{name}/x1"])  # same fp32 op order -> bit exact
    np.testing.assert_array_equal(O.merge(plan, x, "sum"), TC[f"{name}/xsum"])
    plan2 = O.bipartite_soft_matching(TC[f"{name}/metric2"], r, bool(cls), bool(dis))
    np.testing.assert_array_equal(plan2.src_idx, TC[f"{name}/src_idx2"])
    np.testing.assert_array_equal(plan2.dst_idx, TC[f"{name}/dst_idx2"])
    x2, s2 = O.merge_wavg(plan2, x1, s1)
    np.testing.assert_array_equal(s2, TC[f"{name}/s2"])
    np.testing.assert_array_equal(x2, TC[f"{name}/x2"])
    # invariants: token mass and size conservation through two merges
    assert np.all(s2.sum(axis=1) == T)
    np.testing.assert_allclose((s2 * x2).sum(axis=1), x.sum(axis=1), rtol=2e-4, atol=2e-4)


def test_appendix_b_golden_vector():
    metric = np.array([[[1, 0], [1, 0], [0, 1], [3, 4], [1, 0], [0, 2], [4, 3], [-1, 0]]], np.float32)
    x = np.stack([2 * np.arange(8), 2 * np.arange(8) + 1], -1)[None].astype(np.float32)
    plan = O.bipartite_soft_matching(metric, 2)
    np.testing.assert_allclose(plan.scores[0], [[1, .6, 0, -1], [0, .8, 1, 0], [1, .6, 0, -1], [.8, .96, .6, -.8]],
                               atol=1e-6)
    np.testing.assert_array_equal(plan.node_idx[0], [0, 2, 0, 1])
    np.testing.assert_array_equal(plan.edge_idx[0], [2, 1, 0, 3])
    np.testing.assert_array_equal(plan.src_idx[0], [2, 1])
    np.testing.assert_array_equal(plan.dst_idx[0], [0, 2])
    np.testing.assert_array_equal(plan.unm_idx[0], [0, 3])
    x1, s1 = O.merge_wavg(plan, x)
    np.testing.assert_array_equal(x1[0], [[0, 1], [12, 13], [5, 6], [6, 7], [7, 8], [14, 15]])
    np.testing.assert_array_equal(s1[0, :, 0], [1, 1, 2, 1, 2, 1])
    # unmerge: every original row reads its merged row
    um = O.unmerge(plan, x1)
    np.testing.assert_array_equal(um[0, :, 0], [0, 5, 7, 6, 5, 7, 12, 14])
    np.testing.assert_array_equal(O.row_map(plan)[0], [0, 2, 4, 3, 2, 4, 1, 5])


def test_r0_is_identity_and_unmerge_roundtrip():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 9, 4)).astype(np.float32)
    plan = O.bipartite_soft_matching(x, 0)
    x1, s1 = O.merge_wavg(plan, x)
    np.testing.assert_array_equal(x1, x)
    assert np.all(s1 == 1)
    for dis in (False, True):
        plan = O.bipartite_soft_matching(rng.standard_normal((2, 9, 4)).astype(np.float32), 3, False, dis)
        rm = O.row_map(plan)
        # row_map is consistent with merge(): a one-hot "which output row am I in" check
        eye = np.broadcast_to(np.eye(9, dtype=np.float32), (2, 9, 9)).copy()
        merged = O.merge(plan, eye, "sum")  # [2, 6, 9]: merged[b,row,t] = 1 iff token t landed in row
        for b in range(2):
            for t in range(9):
                assert merged[b, rm[b, t], t] == 1 and merged[b, :, t].sum() == 1


@pytest.mark.parametrize("name", [str(n) for n in TS["names"]])
def test_mask_groups_against_reference_golden(name):
    seq = str(TS[f"{name}/seq"])
    T = int(TS[f"{name}/T"])
    gid, pos, allow, ro = O.sequence_groups(seq)
    assert gid.shape[0] == T
    mask = np.unpackbits(TS[f"{name}/mask"], axis=-1)[:, :T].astype(bool)
    np.testing.assert_array_equal(O.dense_mask(gid, pos, gid, pos, allow), mask)
    np.testing.assert_array_equal(ro, TS[f"{name}/readout_idx"])


def test_octo_base_table_matches_survey():
    gid, pos, allow, ro = O.sequence_groups("[TaskDescriptionPrefix{16}] [Image{25};Readout{4}]*2")
    np.testing.assert_array_equal(allow, [[1, 0, 0, 0, 0], [1, 1, 0, 0, 0], [1, 1, 1, 0, 0], [1, 1, 0, 1, 0],
                                          [1, 1, 0, 1, 1]])
    assert gid.shape[0] == 74 and list(ro) == [41, 42, 43, 44, 70, 71, 72, 73]


def test_block_oracle_runs_and_grads_flow():
    import torch
    rng = np.random.default_rng(1)
    C, H, D, Dff, B = 32, 2, 16, 64, 2
    gid, pos, allow, ro = O.sequence_groups("[TaskDescriptionPrefix{4}] [Image{10};Readout{2}]*2")
    T = gid.shape[0]
    params = [O.block_params_to_torch(O.init_block_params(rng, C, H, D, Dff), requires_grad=True) for _ in range(2)]
    x = torch.tensor(rng.standard_normal((B, T, C)).astype(np.float32))
    pe = torch.tensor((rng.standard_normal((1, T, C)) * 0.02).astype(np.float32), requires_grad=True)
    tr = []
    xf, size, origin = O.tome_stack(params, pe, x, gid, pos, allow, num_heads=H, r=3, trace=tr)
    assert xf.shape == (B, T - 6, C) and float(size.sum()) == B * T
    y = torch.tensor(rng.standard_normal((B, len(ro), C)).astype(np.float32))
    loss, _ = O.readout_loss(xf, origin, ro, y)
    loss.backward()
    assert all(t.grad is not None and torch.isfinite(t.grad).all() for p in params for t in p.tensors())
    assert pe.grad is not None


def test_top_k_pruning_matches_reference_goldens():
    """oracle.compute_top_k_tokens vs tests/golden/token_pruning.npz, which oracle/gen_golden.py produced by executing the
    reference's own compute_top_k_tokens (token_compression.py:15-46): kept rows bit for bit, incl. tied and constant scores."""
    g = np.load(os.path.join(GOLD, "token_pruning.npz"))
    for name in [str(n) for n in g["names"]]:
        sets = [tuple(int(v) for v in s) for s in g[f"{name}/sets"]]
        ks = [int(k) for k in g[f"{name}/ks"]]
        kept, ids = O.compute_top_k_tokens(g[f"{name}/emb"], g[f"{name}/imp"], sets, ks)
        np.testing.assert_array_equal(kept, g[f"{name}/kept"], err_msg=name)
        assert ids.shape == (sum(ks),) and len(set(ids.tolist())) == sum(ks)
        if name == "constant":  # the reference's own importance scores are the constant 1/T: the first k of every set survive
            want = np.concatenate([np.arange(s, s + k) for (s, _), k in zip(sets, ks)])
            np.testing.assert_array_equal(ids, want)


def test_action_heads_match_reference_goldens():
    """oracle.continuous_action_head / categorical_action_head / assign_bins / l2_loss / ce_loss against outputs of the
    reference's own action_heads/continuous.py and categorical.py (executed by oracle/gen_golden.py under the shim)."""
    torch = pytest.importorskip("torch")
    g = np.load(os.path.join(GOLD, "action_heads.npz"))
    t = lambda a: torch.tensor(np.asarray(a))  # noqa: E731
    for name in g["continuous"]:
        mx = float(g[f"{name}/max_action"])
        pred = O.continuous_action_head(t(g[f"{name}/readouts"]), t(g[f"{name}/kernel"]), t(g[f"{name}/bias"]), mx)
        assert tuple(pred.shape) == g[f"{name}/pred"].shape                      # [B, 1, A]
        np.testing.assert_allclose(pred.numpy(), g[f"{name}/pred"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(O.l2_loss(pred, t(g[f"{name}/actions"])).numpy(), g[f"{name}/loss"], rtol=1e-4, atol=1e-6)
    for name in g["categorical"]:
        A, bins = (int(v) for v in g[f"{name}/cfg"])
        mx = float(g[f"{name}/max_action"])
        logits = O.categorical_action_head(t(g[f"{name}/readouts"]), t(g[f"{name}/kernel"]), t(g[f"{name}/bias"]), A)
        np.testing.assert_allclose(logits.numpy(), g[f"{name}/logits"], rtol=1e-5, atol=2e-6)
        np.testing.assert_array_equal(O.assign_bins(g[f"{name}/actions"], (-mx, mx), bins), g[f"{name}/target_bin"])   # exact
        np.testing.assert_allclose(O.ce_loss(logits, g[f"{name}/actions"], mx, bins).numpy(), g[f"{name}/loss"], rtol=1e-4, atol=1e-5)
    # the reference's 1-based digitize: the top bin and everything above the range have no class at all
    assert O.assign_bins(np.array([0.999, 1.0, 7.0, -7.0], np.float32), (-1.0, 1.0), 4).tolist() == [4, 5, 5, 0]


def test_diffusion_head_matches_reference_goldens():
    """oracle.cosine_beta_schedule / alpha_hats / octo_denoise / denoise_loss against the reference's own
    cosine_beta_schedule, FourierFeatures, OctoDenoise and MLPBlock (diffusion.py:16-64, attention.py:20-39), executed by
    oracle/gen_golden.py under the shim; the noisy-action and loss lines (:131-142) are restated there on those outputs."""
    torch = pytest.importorskip("torch")
    g = np.load(os.path.join(GOLD, "action_heads.npz"))
    for steps in (8, 32):
        np.testing.assert_array_equal(O.cosine_beta_schedule(steps), g[f"diff_schedule{steps}/betas"])
        np.testing.assert_array_equal(O.alpha_hats(steps), g[f"diff_schedule{steps}/alpha_hats"])
    assert O.cosine_beta_schedule(32)[-1] == np.float32(0.999)             # the clip of :26
    for name in g["diffusion"]:
        p = {k: torch.tensor(g[f"{name}/p/{k}"]) for k in ("fourier_kernel", "tw1", "tb1", "tw2", "tb2", "w1", "b1", "w2", "b2")}
        loss, pred = O.denoise_loss(torch.tensor(g[f"{name}/readouts"]), torch.tensor(g[f"{name}/actions"]),
                                    torch.tensor(g[f"{name}/time"]), torch.tensor(g[f"{name}/noise"]),
                                    O.alpha_hats(int(g[f"{name}/steps"])), p)
        np.testing.assert_allclose(pred.numpy(), g[f"{name}/pred"], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(loss.item(), g[f"{name}/loss"], rtol=1e-5)


# ------------------------------------------------------------------------------------------------ block composition
EB = np.load(os.path.join(GOLD, "encoder_blocks.npz"))


def _eb_tree(name):
    """the Flax parameter tree of an encoder_blocks.npz case, rebuilt from its flattened keys"""
    tree = {}
    pre = f"{name}/params/"
    for k in EB.files:
        if k.startswith(pre):
            node = tree
            parts = k[len(pre):].split("/")
            for p_ in parts[:-1]:
                node = node.setdefault(p_, {})
            node[parts[-1]] = EB[k]
    return tree


def _eb_oracle_params(blk, layer, C):
    import torch
    a = blk["SelfAttention_0"]
    g = lambda t: torch.tensor(np.asarray(t[layer], np.float32))  # noqa: E731
    return O.BlockParams(
        ln1_scale=g(blk["LayerNorm_0"]["scale"]), ln1_bias=g(blk["LayerNorm_0"]["bias"]),
        ln2_scale=g(blk["LayerNorm_1"]["scale"]), ln2_bias=g(blk["LayerNorm_1"]["bias"]),
        wq=g(a["query"]["kernel"]).reshape(C, -1), bq=g(a["query"]["bias"]).reshape(-1),
        wk=g(a["key"]["kernel"]).reshape(C, -1), bk=g(a["key"]["bias"]).reshape(-1),
        wv=g(a["value"]["kernel"]).reshape(C, -1), bv=g(a["value"]["bias"]).reshape(-1),
        wo=g(a["out"]["kernel"]).reshape(-1, C), bo=g(a["out"]["bias"]),
        w1=g(blk["MLPBlock_0"]["Dense_0"]["kernel"]), b1=g(blk["MLPBlock_0"]["Dense_0"]["bias"]),
        w2=g(blk["MLPBlock_0"]["Dense_1"]["kernel"]), b2=g(blk["MLPBlock_0"]["Dense_1"]["bias"]))


def _eb_groups(name, B, T):
    if f"{name}/seq" in EB.files:
        gid, pos, allow, _ = O.sequence_groups(str(EB[f"{name}/seq"]))
    else:
        gid, pos, allow = np.zeros(T, np.uint8), np.arange(T, dtype=np.int32), np.ones((1, 1), np.uint8)
    # the mask the reference itself built (TokenSequence.generate_attention_mask) is what the group table must expand to
    dense = O.dense_mask(np.broadcast_to(gid, (B, T)), np.broadcast_to(pos, (B, T)), np.broadcast_to(gid, (B, T)),
                         np.broadcast_to(pos, (B, T)), allow)
    np.testing.assert_array_equal(dense, EB[f"{name}/mask"][:, 0])
    return gid, pos, allow


@pytest.mark.parametrize("name", [str(n) for n in EB["cases"]])
def test_block_and_stack_follow_the_reference_control_flow(name):
    """oracle.tome_block / tome_stack (r = 0, no size bias: the reference as written) against the outputs of the reference's
    OWN Encoder1DBlock.__call__ and StackedEncoder1DBlock.__call__ (attention.py:41-119), executed by oracle/gen_golden.py
    with the vanilla_decoder.yaml config nodes: which LayerNorm instance feeds what, where the residual adds sit, the MLP
    block's order, the position embedding, the scan over stacked parameters.  fp32 on both sides: 2e-5 relative."""
    import torch
    B, T, C, H, Dff, N, ax = [int(v) for v in EB[f"{name}/meta"]]
    tree = _eb_tree(name)
    blk = tree["ScanEncoder1DBlock_0"]
    gid, pos, allow = _eb_groups(name, B, T)
    x = torch.tensor(EB[f"{name}/x"])
    ln_axis = "seq" if ax == 1 else "feature"
    params = [_eb_oracle_params(blk, l, C) for l in range(N)]
    with torch.no_grad():
        y, size, _ = O.tome_stack(params, torch.tensor(tree["posembed_input"]["pos_embedding"]), x, gid, pos, allow, num_heads=H,
                                  r=0, ln_axis=ln_axis, prop_attn=False)
        g2, p2 = np.broadcast_to(gid, (B, T)).copy(), np.broadcast_to(pos, (B, T)).copy()
        y1, _, _, _ = O.tome_block(params[0], x, torch.ones(B, T, 1), g2, p2, allow, num_heads=H, r=0, ln_axis=ln_axis, prop_attn=False)
    rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))  # noqa: E731
    assert rel(y.numpy(), EB[f"{name}/y"]) <= 2e-5
    assert rel(y1.numpy(), EB[f"{name}/y_block0"]) <= 2e-5
    assert float(size.min()) == 1.0 == float(size.max())


def _golden_tree(Z, prefix):
    """rebuild the nested parameter dict stored under `prefix/...` keys of an .npz"""
    tree = {}
    for k in Z.files:
        if k.startswith(prefix + "/"):
            node = tree
            parts = k[len(prefix) + 1:].split("/")
            for q in parts[:-1]:
                node = node.setdefault(q, {})
            node[parts[-1]] = Z[k]
    return tree


@pytest.mark.parametrize("name", ["two_frames", "nine_patches_one_block", "group_size_one_raw_pixels"])
def test_image_tokenizer_oracle_matches_the_executed_reference(name):
    """oracle.image_to_patches / patch_position_tokens / image_tokenizer_fwd against fixtures made by executing the reference's
    image_tokenizer.py (image_to_patches, encode_patch_position with train=False, ResNetV2Block, ImageTokenizer) under the
    shim: patches and position tokens bit for bit, tokens to 2e-5 (fp32 sums in a different order)."""
    Z = np.load(os.path.join(GOLD, "image_tokenizer.npz"))
    B, N, H, P, Cin, F, G, E, PI, NB, norm = [int(v) for v in Z[f"{name}/meta"]]
    img = Z[f"{name}/image"]
    np.testing.assert_array_equal(O.image_to_patches(img[0, 0].astype(np.float32), P, bool(norm)), Z[f"{name}/patches00"])
    rt, ct = O.patch_position_tokens(H, P, PI)
    np.testing.assert_array_equal(rt, Z[f"{name}/row_tokens"])
    np.testing.assert_array_equal(ct, Z[f"{name}/col_tokens"])
    p = O.image_tokenizer_params_from_flax(_golden_tree(Z, f"{name}/params"), NB)
    got = O.image_tokenizer_fwd(p, img.astype(np.float32), patch_size=P, position_interval=PI, num_groups=G, normalize=bool(norm))
    want = Z[f"{name}/out"]
    assert got.shape == want.shape == (B, N, (H // P) ** 2, E)
    assert np.abs(got - want).max() <= 2e-5 * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("name", ["yaml_axes", "feature_ln", "one_readout"])
def test_attention_pooling_oracle_matches_the_executed_reference(name):
    """oracle.attention_pooling against the output of the reference's own MultiHeadAttentionPooling (attention.py:122-150)
    executed under the shim (tests/golden/attention_pooling.npz): 2e-5 (fp32 sums in a different order)."""
    Z = np.load(os.path.join(GOLD, "attention_pooling.npz"))
    B, n, E, H, Dff, ax = [int(v) for v in Z[f"{name}/meta"]]
    got = O.attention_pooling(Z[f"{name}/x"], _golden_tree(Z, f"{name}/params"), H, ln_axis=ax)
    want = Z[f"{name}/y"]
    assert got.shape == want.shape == (B, 1, E)
    assert np.abs(got - want).max() <= 2e-5 * max(1.0, np.abs(want).max())


def test_image_tokenizer_known_answers_of_the_reference_tests():
    """The reference's own unit tests for this file (tokenizers/images/tests/test_image_tokenizer.py), restated on the oracle and
    on the mirror's host functions: image_to_patches on a 280 x 280 image of sixteen constant 70 x 70 patches in raster order
    (:22-37), encode_patch_position on a 128-pixel image with one-pixel patches and 128 tokens -- row_encoding[123] == 122,
    shape (128^2,) (:40-55) -- and the stochastic case's bound |row_encoding[123] - 122| <= 70 on a 280-pixel image (:58-72)."""
    from multi_modal_transformers_tokenmerge_b200.tokenizers.images import encode_patch_position, image_to_patches_index
    patches = np.ones((16, 70, 70, 3), np.float32) * (np.arange(16, dtype=np.float32) + 1)[:, None, None, None]
    image = patches.reshape(4, 4, 70, 70, 3).transpose(0, 2, 1, 3, 4).reshape(280, 280, 3)      # '(row col) h w c -> (row h) (col w) c'
    np.testing.assert_array_equal(O.image_to_patches(image, 70, normalize=False), patches)
    org = image_to_patches_index(280, 70)
    for k, (y, x) in enumerate(org):
        assert (image[y:y + 70, x:x + 70] == k + 1).all()
    for fn in (lambda: O.patch_position_tokens(128, 1, 128), lambda: encode_patch_position(128, 1, 128, train=False)):
        row, col = fn()
        assert row.shape == col.shape == (128 * 128,) and row.dtype == np.int32
        assert row[123] == 122
    row, col = encode_patch_position(280, 1, 128, train=True, rng=np.random.default_rng(0), images=1)
    assert row.shape == (1, 280 * 280) and abs(int(row[0, 123]) - 122) <= 70
