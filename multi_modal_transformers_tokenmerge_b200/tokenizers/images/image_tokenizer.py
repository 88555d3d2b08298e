"""Mirror of multi_modal_transformers/tokenizers/images/image_tokenizer.py on the sm_100a front-end kernels
(csrc/image_tokenizer.cu).

    ImageTokenizer(image_size, patch_size, normalize, position_interval, rng_collection, embedding_dim,
                   row_position_embedding, col_position_embedding, resnet)            image_tokenizer.py:216-309
        .init(rng, image) -> {"params": {"embedding_function": {Conv_0, GroupNorm_i, Conv_{i+1}, Dense_0},
                                         "image_row_position_embedding": {"embedding"}, "image_col_position_embedding": {...}}}
        .apply(variables, image, train=False) -> [B, N, n_patches, embedding_dim]
    ResNetV2Block(num_blocks, input_conv, input_pool, resnet_norm, resnet_activation, resnet_conv, output_dense)   :148-190
    encode_patch_position(image_size, patch_size, num_tokens, train, rng)            :74-140 (host index arithmetic)

The config nodes are the ones of model_configs/tokenizers/images/gato_resnet.yaml.  What the kernels implement is that
file's structure: a VALID strided input convolution, a stride-1 VALID max pool, GroupNorm -> gelu -> 3x3 SAME convolution
blocks whose output has the pooled tensor's shape (so image_tokenizer.py:167-168's projection of the residual never runs),
Dense.  Anything else raises.  Forward only (the reference trains the tokenizer; its backward is not built here).
Host side: configuration, parameter packing and the position TOKENS (integers); every floating-point operation runs in
the kernels.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict

import numpy as np
import torch

from ... import _lib as L
from ... import ops
from ...attention_blocks._module import Module, _init_name, make_init


def _leaf_ids(tree):
    for k in sorted(tree):
        v = tree[k]
        if isinstance(v, dict):
            yield from _leaf_ids(v)
        else:
            yield id(v)


def _pair(v, what):
    v = list(v) if isinstance(v, (list, tuple)) else [v, v]
    if len(v) != 2 or v[0] != v[1]:
        raise NotImplementedError(f"{what} must be square (got {v})")
    return int(v[0])


def _conv(node, what):
    if node.get("_target_") != "flax.linen.Conv":
        raise ValueError(f"{what}: expected a flax.linen.Conv node, got {node.get('_target_')!r}")
    if not node.get("use_bias", True):
        raise NotImplementedError(f"{what}: use_bias=False is not implemented")
    return dict(features=int(node["features"]), kernel=_pair(node["kernel_size"], f"{what}.kernel_size"),
                stride=_pair(node.get("strides", 1), f"{what}.strides"), padding=str(node.get("padding", "SAME")).upper(),
                kernel_init=_init_name(node.get("kernel_init"), "lecun_normal"), bias_init=_init_name(node.get("bias_init"), "zeros"))


class ResNetV2Block:
    """Configuration holder with the reference's field names (image_tokenizer.py:148-164); ImageTokenizer runs it."""

    def __init__(self, num_blocks: int, input_conv: Dict[str, Any], input_pool: Dict[str, Any], resnet_norm: Dict[str, Any],
                 resnet_activation: Dict[str, Any], resnet_conv: Dict[str, Any], output_dense: Dict[str, Any], **_ignored):
        self.num_blocks = int(num_blocks)
        self.input_conv = _conv(input_conv, "input_conv")
        if self.input_conv["padding"] != "VALID":
            raise NotImplementedError("input_conv: padding VALID only (gato_resnet.yaml)")
        if input_pool.get("_target_") != "flax.linen.max_pool" or str(input_pool.get("padding", "VALID")).upper() != "VALID" \
                or _pair(input_pool.get("strides", 1), "input_pool.strides") != 1:
            raise NotImplementedError("input_pool: flax.linen.max_pool, stride 1, padding VALID only (gato_resnet.yaml)")
        self.pool_window = _pair(input_pool["window_shape"], "input_pool.window_shape")
        if resnet_norm.get("_target_") != "flax.linen.GroupNorm":
            raise NotImplementedError("resnet_norm: flax.linen.GroupNorm only")
        self.num_groups, self.gn_eps = int(resnet_norm.get("num_groups", 32)), float(resnet_norm.get("epsilon", 1e-6))
        if resnet_activation.get("_target_") not in ("flax.linen.gelu", "jax.nn.gelu") or resnet_activation.get("approximate", True) is not True:
            raise NotImplementedError("resnet_activation: flax.linen.gelu (tanh approximation) only")
        self.resnet_conv = _conv(resnet_conv, "resnet_conv")
        if (self.resnet_conv["kernel"], self.resnet_conv["stride"], self.resnet_conv["padding"]) != (3, 1, "SAME") \
                or self.resnet_conv["features"] != self.input_conv["features"]:
            raise NotImplementedError("resnet_conv: 3x3, stride 1, SAME, as many features as input_conv (gato_resnet.yaml); the "
                                      "projected-residual branch of image_tokenizer.py:167-168 is not implemented")
        if output_dense.get("_target_") != "flax.linen.Dense" or not output_dense.get("use_bias", True):
            raise NotImplementedError("output_dense: flax.linen.Dense with bias")
        self.dense_features = int(output_dense["features"])
        self.dense_kernel_init = _init_name(output_dense.get("kernel_init"), "lecun_normal")
        self.dense_bias_init = _init_name(output_dense.get("bias_init"), "zeros")


def image_to_patches_index(image_size: int, patch_size: int) -> np.ndarray:
    """Pixel origin (y, x) of patch k in the (h w) order of image_to_patches (image_tokenizer.py:54-62) -> int32 [n, 2].
    The kernels extract patches by this index arithmetic; no patch tensor is materialised."""
    if image_size % patch_size:
        raise ValueError("image_size must be a multiple of patch_size")
    ppd = image_size // patch_size
    k = np.arange(ppd * ppd)
    return np.stack([(k // ppd) * patch_size, (k % ppd) * patch_size], axis=1).astype(np.int32)


def encode_patch_position(image_size: int, patch_size: int, num_tokens: int, train: bool = False, rng=None, images: int = 1):
    """image_tokenizer.py:74-140.  Interval edges of patch k's "row" (index k % patches_per_dim, :94) and "col"
    (k // patches_per_dim, :95) are normalised by the image size, scaled to num_tokens - 1 and floored in fp32 (:100);
    evaluation takes the floor-divided midpoint (:112-113), training draws uniformly from [start, stop) per image and patch
    (:103-108; numpy's generator instead of jax.random, so the draws differ from the reference's stream).
    Returns int32 (row_tokens, col_tokens), [n_patches] for evaluation, [images, n_patches] for training."""
    ppd = image_size // patch_size
    f = np.float32
    edges = np.arange(0, image_size + patch_size, patch_size)
    q = np.floor((edges.astype(f) / f(image_size)) * f(num_tokens - 1)).astype(f)
    k = np.arange(ppd * ppd)
    if not train:
        mid = np.floor_divide(q[:-1] + q[1:], f(2)).astype(np.int32)
        return mid[k % ppd].astype(np.int32), mid[k // ppd].astype(np.int32)
    if rng is None:
        raise ValueError("train=True draws the position tokens: pass rng (the reference's rng_collection key)")
    lo, hi = q[:-1].astype(np.int64), q[1:].astype(np.int64)
    hi = np.maximum(hi, lo + 1)                           # jax.random.randint returns minval for an empty range
    row = rng.integers(lo[k % ppd], hi[k % ppd], size=(images, k.size))
    col = rng.integers(lo[k // ppd], hi[k // ppd], size=(images, k.size))
    return row.astype(np.int32), col.astype(np.int32)


class ImageTokenizer(Module):
    def __init__(self, image_size, patch_size: int, normalize: bool, position_interval: int, rng_collection: str, embedding_dim: int,
                 row_position_embedding: Dict[str, Any], col_position_embedding: Dict[str, Any], resnet: Dict[str, Any],
                 out_dtype: torch.dtype = torch.bfloat16, chunk_rows: int = 0):
        self.image_size = tuple(int(v) for v in image_size)
        if len(self.image_size) != 3 or self.image_size[0] != self.image_size[1]:
            raise NotImplementedError("image_size must be (H, H, C): encode_patch_position assumes square images (:83-85)")
        self.patch_size, self.normalize, self.position_interval = int(patch_size), bool(normalize), int(position_interval)
        self.rng_collection, self.embedding_dim, self.out_dtype = rng_collection, int(embedding_dim), out_dtype
        self.chunk_rows = int(chunk_rows)
        for nm, node in (("row", row_position_embedding), ("col", col_position_embedding)):
            if node.get("_target_") != "flax.linen.Embed" or int(node["num_embeddings"]) != self.position_interval \
                    or int(node["features"]) != self.embedding_dim:
                raise ValueError(f"{nm}_position_embedding must be flax.linen.Embed(num_embeddings=position_interval, features=embedding_dim)")
        self.row_name = row_position_embedding.get("name", "row_embeddings")
        self.col_name = col_position_embedding.get("name", "col_embeddings")
        r = dict(resnet)
        r.pop("_target_", None)
        r.pop("_recursive_", None)
        self.resnet = ResNetV2Block(**r)
        if self.resnet.dense_features != self.embedding_dim:
            raise ValueError("output_dense.features must equal embedding_dim (the position embeddings are added to it, :306)")
        self._packed = None

    # ---- descriptor and parameter packing -------------------------------------------------------------------
    def _desc(self, batch: int, n_images: int, image_dtype: torch.dtype, token_rows: int) -> "L.ImageTokenizerDesc":
        r = self.resnet
        return L.ImageTokenizerDesc(batch=batch, n_images=n_images, image_size=self.image_size[0], channels_in=self.image_size[2],
                                    image_dtype=L.TOME_U8 if image_dtype == torch.uint8 else L.TOME_F32, normalize=int(self.normalize),
                                    patch_size=self.patch_size, conv_kernel=r.input_conv["kernel"], conv_stride=r.input_conv["stride"],
                                    features=r.input_conv["features"], pool_window=r.pool_window, num_blocks=r.num_blocks,
                                    num_groups=r.num_groups, gn_eps=r.gn_eps, embed_dim=self.embedding_dim,
                                    position_interval=self.position_interval, token_rows=token_rows,
                                    out_dtype=L.TOME_BF16 if self.out_dtype == torch.bfloat16 else L.TOME_F32,
                                    chunk_rows=self.chunk_rows)

    def _geometry(self):
        r = self.resnet
        o1 = (self.patch_size - r.input_conv["kernel"]) // r.input_conv["stride"] + 1
        return o1, o1 - (r.pool_window - 1)

    def _init(self, rng, image=None):
        r, F, E, cin = self.resnet, self.resnet.input_conv["features"], self.embedding_dim, self.image_size[2]
        k0 = r.input_conv["kernel"]
        _, o2 = self._geometry()
        ef = {"Conv_0": {"kernel": make_init(r.input_conv["kernel_init"])(rng, (k0, k0, cin, F), k0 * k0 * cin, F),
                         "bias": make_init(r.input_conv["bias_init"])(rng, (F,))}}
        for i in range(r.num_blocks):
            ef[f"GroupNorm_{i}"] = {"scale": np.ones(F, np.float32), "bias": np.zeros(F, np.float32)}
            ef[f"Conv_{i + 1}"] = {"kernel": make_init(r.resnet_conv["kernel_init"])(rng, (3, 3, F, F), 9 * F, F),
                                   "bias": make_init(r.resnet_conv["bias_init"])(rng, (F,))}
        kd = o2 * o2 * F
        ef["Dense_0"] = {"kernel": make_init(r.dense_kernel_init)(rng, (kd, E), kd, E), "bias": make_init(r.dense_bias_init)(rng, (E,))}
        emb = lambda: {"embedding": (rng.standard_normal((self.position_interval, E)) / np.sqrt(E)).astype(np.float32)}  # noqa: E731
        return {"embedding_function": ef, self.row_name: emb(), self.col_name: emb()}

    def pack_params(self, params, device="cuda") -> torch.Tensor:
        """Flax parameter tree -> the flat fp32 vector of include/tome_b200.h section 7b (kernels keep Flax's [.., in, out])."""
        d = self._desc(1, 1, torch.uint8, 1)
        lib = L.lib()
        n = int(lib.tome_image_tokenizer_param_count(C.byref(d)))
        if n < 0:
            L.check(1)
        flat = np.zeros(n, np.float32)

        def put(which, a, shape):
            a = np.asarray(a.detach().cpu() if isinstance(a, torch.Tensor) else a, np.float32)
            if tuple(a.shape) != tuple(shape):
                raise ValueError(f"parameter {which}: shape {tuple(a.shape)} != {tuple(shape)}")
            o = int(lib.tome_image_tokenizer_param_offset(C.byref(d), which))
            flat[o:o + a.size] = a.ravel()

        r, F, E, cin = self.resnet, self.resnet.input_conv["features"], self.embedding_dim, self.image_size[2]
        k0 = r.input_conv["kernel"]
        _, o2 = self._geometry()
        ef = params["embedding_function"]
        put(L.IT_CONV0_KERNEL, ef["Conv_0"]["kernel"], (k0, k0, cin, F))
        put(L.IT_CONV0_BIAS, ef["Conv_0"]["bias"], (F,))
        for i in range(r.num_blocks):
            put(L.IT_BLOCK0 + 4 * i + 0, ef[f"GroupNorm_{i}"]["scale"], (F,))
            put(L.IT_BLOCK0 + 4 * i + 1, ef[f"GroupNorm_{i}"]["bias"], (F,))
            put(L.IT_BLOCK0 + 4 * i + 2, ef[f"Conv_{i + 1}"]["kernel"], (3, 3, F, F))
            put(L.IT_BLOCK0 + 4 * i + 3, ef[f"Conv_{i + 1}"]["bias"], (F,))
        put(L.IT_DENSE_KERNEL, ef["Dense_0"]["kernel"], (o2 * o2 * F, E))
        put(L.IT_DENSE_BIAS, ef["Dense_0"]["bias"], (E,))
        put(L.IT_ROW_EMBED, params[self.row_name]["embedding"], (self.position_interval, E))
        put(L.IT_COL_EMBED, params[self.col_name]["embedding"], (self.position_interval, E))
        return torch.from_numpy(flat).to(device)

    # ---- forward ----------------------------------------------------------------------------------------------
    def _apply(self, params, image, train: bool = True, dropout_rng=None, patch_rng=None):
        if not isinstance(image, torch.Tensor):
            image = torch.as_tensor(np.asarray(image))
        if not image.is_cuda:
            raise RuntimeError("ImageTokenizer runs on CUDA (sm_100a) only: got a CPU tensor.  There is no CPU fallback.")
        if image.dim() != 5 or tuple(image.shape[-3:]) != self.image_size:
            raise ValueError(f"Input image is not the correct size: {tuple(image.shape)} vs [B, N, {self.image_size}] (:240-243)")
        if image.dtype not in (torch.uint8, torch.float32):
            image = image.to(torch.float32)
        B, N = int(image.shape[0]), int(image.shape[1])
        if train:
            row, col = encode_patch_position(self.image_size[0], self.patch_size, self.position_interval, True,
                                             np.random.default_rng(patch_rng) if not isinstance(patch_rng, np.random.Generator) else patch_rng,
                                             images=B * N)
        else:
            row, col = encode_patch_position(self.image_size[0], self.patch_size, self.position_interval, False)
        # the packed device copy is cached per parameter-tree OBJECT and per leaf object: replacing a leaf (what an optimiser
        # update of a Flax tree does) repacks; arrays mutated in place are not noticed -- call clear_cache() after doing that
        key = (id(params),) + (() if isinstance(params, torch.Tensor) else tuple(_leaf_ids(params)))
        if self._packed is None or self._packed[0] != key:
            flat = params if isinstance(params, torch.Tensor) else self.pack_params(params, image.device)
            self._packed = (key, flat, flat.to(torch.bfloat16))
        _, flat, flat16 = self._packed
        d = self._desc(B, N, image.dtype, B * N if train else 1)
        dev = image.device
        return ops.image_tokenizer_fwd(image.contiguous(), flat, d, torch.from_numpy(np.ascontiguousarray(row)).to(dev),
                                       torch.from_numpy(np.ascontiguousarray(col)).to(dev), params_bf16=flat16)

    def clear_cache(self) -> None:
        """Drop the packed device copy of the parameters (needed only after mutating parameter arrays in place)."""
        self._packed = None

    def apply(self, variables, image, train: bool = True, rngs=None, **kw):
        if rngs is not None and self.rng_collection in rngs:
            kw.setdefault("patch_rng", int(np.asarray(rngs[self.rng_collection]).ravel()[-1]))
        return super().apply(variables, image, train=train, **kw)
