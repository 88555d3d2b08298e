"""Token-sequence grammar -> what the CUDA attention kernels consume: a per-token GROUP ID, a position inside the
group, and a G x G allow table, instead of a dense [B, H, T, T] boolean mask.

Mirror of the reference's multi_modal_transformers/tokenizers/token_sequencer.py (same class / method names):
  TokenSet, Text, TaskDescriptionPrefix, Image, Readout  token_sequencer.py:19-183  (attention rules)
  TokenSequence                                          token_sequencer.py:186-340 (grammar "[A{n};B{m}]*k")
  TokenEmbeddings                                        token_sequencer.py:342-346
Host-side integer logic only (numpy); `generate_attention_mask` still returns the dense mask for API compatibility
and for tests, but the fast path never builds it (octo.py:66-68,119 materialise 883 MB of it at the C3 shape).

Allow-table codes: 0 = masked, 1 = visible, 2 = visible iff pos_key <= pos_query (the causal rule inside a Text set).
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

MASKED, VISIBLE, CAUSAL = 0, 1, 2


class TokenSet:
    """A run of `num_tokens` tokens of one modality at one timestep (token_sequencer.py:19-52)."""

    modality = ""

    def __init__(self, num_tokens: int, timestep: int):
        self.num_tokens = int(num_tokens)
        self.timestep = int(timestep)

    # rule(query=self, key) -> allow code; subclasses implement the table of SURVEY.md A.6
    def rule(self, key: "TokenSet") -> int:
        raise NotImplementedError

    def attention_rule(self, token_sequence: Sequence["TokenSet"]) -> np.ndarray:
        """Dense rows of the mask for this set's queries against every key set (token_sequencer.py:84-91 etc.)."""
        cols = []
        for key in token_sequence:
            code = self.rule(key)
            if code == CAUSAL:
                blk = np.tril(np.ones((self.num_tokens, key.num_tokens), dtype=np.int32))
            else:
                blk = np.full((self.num_tokens, key.num_tokens), 1 if code == VISIBLE else 0, dtype=np.int32)
            cols.append(blk)
        return np.hstack(cols)


class Text(TokenSet):
    """Causal inside its own set, sees earlier-or-same timesteps of text / images, never readouts (:55-91)."""

    modality = "text"

    def rule(self, key):
        if key.timestep == self.timestep and isinstance(key, Text):
            return CAUSAL
        if isinstance(key, Readout):
            return MASKED
        return VISIBLE if key.timestep <= self.timestep else MASKED


class TaskDescriptionPrefix(Text):
    """Bidirectional inside the prefix, sees nothing else (:94-113)."""

    def rule(self, key):
        return VISIBLE if (key.timestep == self.timestep and isinstance(key, type(self))) else MASKED


class Image(TokenSet):
    """Bidirectional inside its frame, sees earlier-or-same timesteps of text / images, never readouts (:116-148)."""

    modality = "images"

    def rule(self, key):
        if key.timestep == self.timestep and isinstance(key, type(self)):
            return VISIBLE
        if isinstance(key, Readout):
            return MASKED
        return VISIBLE if key.timestep <= self.timestep else MASKED


class Readout(TokenSet):
    """Sees its own set and earlier-or-same text / images; no other readouts; nobody sees it (:151-183)."""

    modality = "readouts"

    def rule(self, key):
        if key.timestep == self.timestep and isinstance(key, type(self)):
            return VISIBLE
        if isinstance(key, Readout):
            return MASKED
        return VISIBLE if key.timestep <= self.timestep else MASKED


_SETS = {"Text": Text, "TaskDescriptionPrefix": TaskDescriptionPrefix, "Image": Image, "Readout": Readout}


@dataclass
class TokenEmbeddings:
    """token_sequencer.py:342-346."""

    text: Optional[object] = None
    images: Optional[object] = None
    readouts: Optional[object] = None


class TokenSequence:
    def __init__(self, token_sequence: str, token_compression_sequence: Optional[str] = None):
        self.token_sequence_str = token_sequence
        self.token_compression_sequence_str = token_compression_sequence
        self.token_sequence: List[TokenSet] = self._parse()

    # ---------------------------------------------------------------- grammar (token_sequencer.py:199-253)
    def _parse(self, layer: int = 0) -> List[TokenSet]:
        blocks = re.findall(r"\[(.*?)\]", self.token_sequence_str)
        repeats = []
        for rep in re.findall(r"(?<=\])(.*?)(?=\[|$)", self.token_sequence_str):
            repeats.append(1 if rep.strip() == "" else int(re.findall(r"\*(\d+)", rep)[0]))
        if len(repeats) != len(blocks):
            raise ValueError(f"malformed token sequence {self.token_sequence_str!r}")
        comp = (re.findall(r"\[(.*?)\]", self.token_compression_sequence_str)
                if self.token_compression_sequence_str is not None else [None] * len(blocks))
        seq, timestep = [], 0
        for block, cblock, rep in zip(blocks, comp, repeats):
            groups = block.split(";")
            cgroups = cblock.split(";") if cblock is not None else [None] * len(groups)
            for _ in range(rep):
                for grp, cgrp in zip(groups, cgroups):
                    name = re.search(r"^(.*?)\{", grp).group(1).strip()
                    if name not in _SETS:
                        raise ValueError(f"unknown token set {name!r} (known: {sorted(_SETS)})")
                    n = int(re.search(r"\d+", grp).group())
                    if cgrp is not None:  # per-layer pruning grammar (:222-238)
                        n -= layer * int(re.search(r"\d+", cgrp).group())
                    seq.append(_SETS[name](n, timestep))
                timestep += 1
        return seq

    def generate_layer_token_sequence(self, layer: int) -> List[TokenSet]:
        return self._parse(layer=layer)

    @property
    def num_tokens(self) -> int:
        return sum(s.num_tokens for s in self.token_sequence)

    # ---------------------------------------------------------------- what the kernels consume
    def group_ids(self) -> Tuple[np.ndarray, np.ndarray]:
        """(gid uint8 [T], pos int32 [T]): group = index of the token's set, pos = index inside the set."""
        gid = np.concatenate([np.full(s.num_tokens, g, np.uint8) for g, s in enumerate(self.token_sequence)])
        pos = np.concatenate([np.arange(s.num_tokens, dtype=np.int32) for s in self.token_sequence])
        return gid, pos

    def layer_group_ids(self, layer: int) -> Tuple[np.ndarray, np.ndarray]:
        """group_ids() of the compression grammar at `layer` (the rows of generate_attention_mask(layer=layer), :222-238,
        313-321, as group ids / positions): what a pruning stack passes as tome_stack_io_t.layer_gid / layer_pos."""
        sets = self._parse(layer=layer)
        gid = np.concatenate([np.full(s.num_tokens, g, np.uint8) for g, s in enumerate(sets)])
        pos = np.concatenate([np.arange(s.num_tokens, dtype=np.int32) for s in sets])
        return gid, pos

    def prune_sets(self) -> List[Tuple[int, int]]:
        """[(tokens at layer 0, tokens dropped by every layer)] per token set, in sequence order: tome_stack_cfg_t.prune_set_n /
        prune_set_c (the grammar's `num_tokens - layer * num_compressed_tokens`, :236)."""
        if self.token_compression_sequence_str is None:
            raise ValueError("no token_compression_sequence was given")
        l0, l1 = self._parse(layer=0), self._parse(layer=1)
        return [(a.num_tokens, a.num_tokens - b.num_tokens) for a, b in zip(l0, l1)]

    def allow_table(self) -> np.ndarray:
        """uint8 [G, G]: rule code of (query group, key group)."""
        sets = self.token_sequence
        if len(sets) > 32:
            raise ValueError(f"{len(sets)} token sets: the attention kernels support at most 32 groups")
        return np.array([[q.rule(k) for k in sets] for q in sets], dtype=np.uint8)

    # ---------------------------------------------------------------- reference API (dense mask, indices, assembly)
    def generate_attention_mask(self, repeats: int = 1, layer: Optional[int] = None) -> np.ndarray:
        """Dense bool [repeats, Tq, Tk] (token_sequencer.py:313-321).  Keys always come from the uncompressed
        sequence, as in the reference."""
        qsets = self._parse(layer=layer or 0)
        mask = np.vstack([s.attention_rule(self.token_sequence) for s in qsets]).astype(bool)
        return np.broadcast_to(mask, (repeats,) + mask.shape).copy()

    def get_modality_idx(self, modality: str) -> np.ndarray:
        """Sequence positions of all tokens of one modality (token_sequencer.py:323-334)."""
        idx, cur = [], 0
        for s in self.token_sequence:
            if s.modality == modality:
                idx.append(np.arange(cur, cur + s.num_tokens))
            cur += s.num_tokens
        return np.concatenate(idx).astype(np.int32) if idx else np.zeros(0, np.int32)

    def assemble_embeddings(self, embeddings: TokenEmbeddings):
        """Concatenate per-modality embeddings [B, n, E] in grammar order (token_sequencer.py:255-270).  Works on
        torch tensors or numpy arrays (a pure re-ordering; no arithmetic)."""
        cursor = {"text": 0, "images": 0, "readouts": 0}
        parts = []
        for s in self.token_sequence:
            src = getattr(embeddings, s.modality)
            a = cursor[s.modality]
            parts.append(src[:, a: a + s.num_tokens])
            cursor[s.modality] = a + s.num_tokens
        if hasattr(parts[0], "is_cuda") or type(parts[0]).__module__.startswith("torch"):
            import torch

            return torch.cat(parts, dim=1)
        return np.concatenate(parts, axis=1)


def sequence_groups(seq: str):
    """-> (gid uint8 [T], pos int32 [T], allow uint8 [G, G], readout_idx int32 [n])."""
    ts = TokenSequence(seq)
    gid, pos = ts.group_ids()
    return gid, pos, ts.allow_table(), ts.get_modality_idx("readouts")
