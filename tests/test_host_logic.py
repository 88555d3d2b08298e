"""CPU-only: the product's host-side mirror of tokenizers/token_sequencer.py against the goldens produced by executing
the reference (tests/golden/token_sequencer.npz, made by oracle/gen_golden.py) and against the oracle."""
import os

import numpy as np
import pytest

from multi_modal_transformers_tokenmerge_b200.tokenizers.token_sequencer import (TokenEmbeddings, TokenSequence,
                                                                                 sequence_groups)
from oracle import tome_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TS = np.load(os.path.join(GOLD, "token_sequencer.npz"))


@pytest.mark.parametrize("name", [str(n) for n in TS["names"]])
def test_token_sequence_against_reference_golden(name):
    seq = str(TS[f"{name}/seq"])
    T = int(TS[f"{name}/T"])
    ts = TokenSequence(seq)
    assert ts.num_tokens == T
    mask = np.unpackbits(TS[f"{name}/mask"], axis=-1)[:, :T].astype(bool)
    np.testing.assert_array_equal(ts.generate_attention_mask(repeats=2, layer=0)[1], mask)
    np.testing.assert_array_equal(ts.get_modality_idx("readouts"), TS[f"{name}/readout_idx"])
    # the compact form the kernels consume expands to the same dense mask
    gid, pos = ts.group_ids()
    allow = ts.allow_table()
    dense = (allow[gid[:, None], gid[None, :]] == 1) | ((allow[gid[:, None], gid[None, :]] == 2) & (pos[None, :] <= pos[:, None]))
    np.testing.assert_array_equal(dense, mask)
    for a, b in zip(sequence_groups(seq), O.sequence_groups(seq)):
        np.testing.assert_array_equal(a, b)


def test_assemble_embeddings_and_compressed_grammar():
    ts = TokenSequence("[TaskDescriptionPrefix{2}] [Image{3};Readout{1}]*2")
    emb = TokenEmbeddings(text=np.arange(2)[None, :, None] + 100.0, images=np.arange(6)[None, :, None] + 200.0,
                          readouts=np.arange(2)[None, :, None] + 300.0)
    out = ts.assemble_embeddings(emb)[0, :, 0]
    np.testing.assert_array_equal(out, [100, 101, 200, 201, 202, 300, 203, 204, 205, 301])
    tc = TokenSequence("[TaskDescriptionPrefix{4}] [Image{10};Readout{2}]*2", "[TaskDescriptionPrefix{0}] [Image{2};Readout{0}]*2")
    assert [s.num_tokens for s in tc.generate_layer_token_sequence(3)] == [4, 4, 2, 4, 2]
    assert tc.generate_attention_mask(1, layer=3).shape == (1, 16, 28)  # rectangular, as in the reference
    with pytest.raises(ValueError):
        TokenSequence("[Bogus{3}]")


def test_image_patch_position_tokens_match_the_executed_reference():
    """tokenizers/images mirror: evaluation-mode position tokens against the reference's own encode_patch_position
    (tests/golden/image_tokenizer.npz), the patch origin table against image_to_patches' (h w) order, and the training-mode
    draws inside their quantised intervals."""
    import os
    from multi_modal_transformers_tokenmerge_b200.tokenizers.images import encode_patch_position, image_to_patches_index
    Z = np.load(os.path.join(os.path.dirname(__file__), "golden", "image_tokenizer.npz"))
    for name in Z["cases"]:
        B, N, H, P, Cin, F, G, E, PI, NB, norm = [int(v) for v in Z[f"{name}/meta"]]
        rt, ct = encode_patch_position(H, P, PI, train=False)
        np.testing.assert_array_equal(rt, Z[f"{name}/row_tokens"])
        np.testing.assert_array_equal(ct, Z[f"{name}/col_tokens"])
        org = image_to_patches_index(H, P)
        img = Z[f"{name}/image"][0, 0].astype(np.float32)
        want = Z[f"{name}/patches00"]
        for k, (y, x) in enumerate(org):
            px = img[y:y + P, x:x + P]
            if norm:
                px = np.float32(2) * (px / np.float32(255)) - np.float32(1)
            np.testing.assert_array_equal(px, want[k])
        r2, c2 = encode_patch_position(H, P, PI, train=True, rng=np.random.default_rng(3), images=5)
        ppd = H // P
        q = np.floor((np.arange(0, H + P, P, dtype=np.float32) / np.float32(H)) * np.float32(PI - 1)).astype(np.int64)
        k = np.arange(ppd * ppd)
        assert r2.shape == c2.shape == (5, ppd * ppd)
        assert (r2 >= q[k % ppd]).all() and (r2 < np.maximum(q[k % ppd + 1], q[k % ppd] + 1)).all()
        assert (c2 >= q[k // ppd]).all() and (c2 < np.maximum(q[k // ppd + 1], q[k // ppd] + 1)).all()
