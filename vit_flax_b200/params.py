"""The ``ViT`` params pytree: names, shapes, initialisers, flattening.

Mirrors what ``ViT.init`` produces in the reference (vit.py:187-191): Flax
compact auto-naming puts ``Attention_l`` / ``FeedForward_l`` / ``PreNorm_k`` as
siblings under ``Transformer_0`` (modules built at vit.py:104-105 are bound to
the Transformer scope), ``Residual`` / ``IdentityLayer`` / ``Dropout`` own no
params.  Leaves are looked up BY NAME, never by flattened position (jax sorts
keys lexicographically: ``Attention_10`` < ``Attention_2``).
"""
from __future__ import annotations

from collections.abc import Mapping
from typing import Dict, Iterator, List, Tuple

import numpy as np

DIM_HEAD = 64  # vit.py:123


def pair(t):
    """vit.py:15-16."""
    return t if isinstance(t, tuple) else (t, t)


def geometry(image_size, patch_size):
    ih, iw = pair(image_size)
    ph, pw = pair(patch_size)
    assert ih % ph == 0, "Image dimensions must be divisible by the patch size."   # vit.py:133
    assert iw % pw == 0, "Image dimensions must be divisible by the patch size."   # vit.py:134
    num_patches = (ih // ph) * (iw // pw)                                          # vit.py:136
    return ih, iw, ph, pw, num_patches


def param_specs(*, image_size, patch_size, num_classes, dim, depth, heads, mlp_dim,
                channels=3) -> List[Tuple[Tuple[str, ...], Tuple[int, ...], str]]:
    """Ordered ``(path, shape, initialiser)`` for every leaf.

    Initialisers: ``zeros`` (pos_embedding, cls -- vit.py:142,144 -- and every
    Dense/LayerNorm bias), ``ones`` (LayerNorm scale), ``lecun_normal`` (Dense
    kernels, flax default).
    """
    _, _, ph, pw, n = geometry(image_size, patch_size)
    inner = DIM_HEAD * heads
    project_out = not (heads == 1 and DIM_HEAD == dim)      # vit.py:65
    k0 = ph * pw * channels
    specs = [
        (("pos_embedding",), (1, n + 1, dim), "zeros"),
        (("cls",), (1, 1, dim), "zeros"),
        (("Dense_0", "kernel"), (k0, dim), "lecun_normal"),
        (("Dense_0", "bias"), (dim,), "zeros"),
    ]
    t = "Transformer_0"
    for l in range(depth):
        specs.append(((t, f"Attention_{l}", "Dense_0", "kernel"), (dim, 3 * inner), "lecun_normal"))
        if project_out:
            specs.append(((t, f"Attention_{l}", "Dense_1", "kernel"), (inner, dim), "lecun_normal"))
            specs.append(((t, f"Attention_{l}", "Dense_1", "bias"), (dim,), "zeros"))
        specs.append(((t, f"FeedForward_{l}", "Dense_0", "kernel"), (dim, mlp_dim), "lecun_normal"))
        specs.append(((t, f"FeedForward_{l}", "Dense_0", "bias"), (mlp_dim,), "zeros"))
        specs.append(((t, f"FeedForward_{l}", "Dense_1", "kernel"), (mlp_dim, dim), "lecun_normal"))
        specs.append(((t, f"FeedForward_{l}", "Dense_1", "bias"), (dim,), "zeros"))
        for k in (2 * l, 2 * l + 1):
            specs.append(((t, f"PreNorm_{k}", "LayerNorm_0", "scale"), (dim,), "ones"))
            specs.append(((t, f"PreNorm_{k}", "LayerNorm_0", "bias"), (dim,), "zeros"))
    specs += [
        (("LayerNorm_0", "scale"), (dim,), "ones"),
        (("LayerNorm_0", "bias"), (dim,), "zeros"),
        (("Dense_1", "kernel"), (dim, num_classes), "lecun_normal"),
        (("Dense_1", "bias"), (num_classes,), "zeros"),
    ]
    return specs


def count_params(**cfg) -> int:
    return int(sum(int(np.prod(s)) for _, s, _ in param_specs(**cfg)))


def _lecun_normal(rng: np.random.Generator, shape) -> np.ndarray:
    """``jax.nn.initializers.lecun_normal``: truncated normal on [-2, 2] scaled so
    the variance is 1/fan_in (fan_in = shape[0] for a Dense kernel)."""
    std = np.sqrt(1.0 / shape[0]) / 0.87962566103423978
    x = rng.standard_normal(shape)
    bad = np.abs(x) > 2.0
    while bad.any():
        x[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(x) > 2.0
    return (x * std).astype(np.float32)


def init_params(seed: int = 0, **cfg) -> Dict:
    """Same tree / shapes / dtypes / distributions as the reference ``init``.
    Values are not bit-identical to JAX's threefry stream (JAX is not available
    here); SURVEY.md section 8b explains why that is not part of the contract."""
    rng = np.random.default_rng(seed)
    tree: Dict = {}
    for path, shape, kind in param_specs(**cfg):
        if kind == "zeros":
            leaf = np.zeros(shape, np.float32)
        elif kind == "ones":
            leaf = np.ones(shape, np.float32)
        else:
            leaf = _lecun_normal(rng, shape)
        node = tree
        for k in path[:-1]:
            node = node.setdefault(k, {})
        node[path[-1]] = leaf
    return {"params": tree}


def perturb_params(variables: Dict, seed: int = 7, std: float = 0.02, head_gain: float = 1.0) -> Dict:
    """Copy with every zeros/ones-initialised leaf += N(0, std): reference-init
    params never exercise cls / pos / bias / LN-affine paths.
    ``head_gain`` sharpens the classifier kernel (top-1 margins)."""
    rng = np.random.default_rng(seed)

    def walk(node, path):
        out = {}
        for k, v in node.items():
            if isinstance(v, Mapping):
                out[k] = walk(v, path + (k,))
            else:
                a = np.array(v, dtype=np.float32, copy=True)
                if k in ("bias", "scale", "cls", "pos_embedding"):
                    a = a + rng.standard_normal(a.shape).astype(np.float32) * std
                if path == ("Dense_1",) and k == "kernel":
                    a = a * np.float32(head_gain)
                out[k] = a
        return out

    p = variables["params"] if "params" in variables else variables
    return {"params": walk(p, ())}


def iter_leaves(tree, prefix: Tuple[str, ...] = ()) -> Iterator[Tuple[str, object]]:
    """Depth-first ``('A/B/leaf', array)`` over any Mapping tree (dict, FrozenDict...)."""
    for k, v in tree.items():
        if isinstance(v, Mapping) or (hasattr(v, "items") and not hasattr(v, "shape")):
            yield from iter_leaves(v, prefix + (str(k),))
        else:
            yield "/".join(prefix + (str(k),)), v


def flatten_params(variables) -> Dict[str, object]:
    tree = variables["params"] if "params" in variables else variables
    return dict(iter_leaves(tree))


def leaf_to_numpy(a) -> np.ndarray:
    """numpy / torch / jax / anything with ``__array__`` or ``__dlpack__`` -> C-contiguous fp32."""
    if isinstance(a, np.ndarray):
        out = a
    elif hasattr(a, "detach") and hasattr(a, "cpu"):        # torch.Tensor
        out = a.detach().cpu().numpy()
    else:
        try:
            out = np.asarray(a)
        except Exception:                                   # pragma: no cover - exotic leaves
            out = np.from_dlpack(a)
    return np.ascontiguousarray(out, dtype=np.float32)
