"""The reference's Python surface (vit.py:114-125, 187-192) as mirrored by vit_flax_b200.ViT."""
import dataclasses
from collections import OrderedDict
from types import MappingProxyType

import numpy as np
import pytest
import torch

from vit_flax_b200 import ViT, flatten_params
from vit_flax_b200.params import leaf_to_numpy, param_specs
from vit_flax_b200.vit import tree_shapes
from _util import C1, TINY


def test_constructor_fields_order_and_defaults():
    names = [f.name for f in dataclasses.fields(ViT)]
    assert names == ["image_size", "patch_size", "num_classes", "dim", "depth", "heads", "mlp_dim",
                     "pool", "dropout", "emb_dropout"]
    v = ViT(256, 32, 1000, 1024, 6, 16, 2048)                 # positional, like a flax dataclass
    assert (v.pool, v.dropout, v.emb_dropout, v.dim_head) == ("cls", 0.0, 0.0, 64)
    with pytest.raises(TypeError):
        ViT(dim_head=32, **C1)                                # vit.py:123: not a field
    with pytest.raises(dataclasses.FrozenInstanceError):
        v.dim = 3
    assert hash(v) == hash(ViT(**C1))


def test_init_tree_matches_flax_naming():
    v = ViT(**C1)
    img = np.zeros((1, 256, 256, 3), np.float32)
    variables = v.init({"params": 1, "dropout": 2, "emb_dropout": 3}, img)   # vit.py:187-191
    assert list(variables) == ["params"]
    p = variables["params"]
    assert set(p) == {"pos_embedding", "cls", "Dense_0", "Transformer_0", "LayerNorm_0", "Dense_1"}
    t = p["Transformer_0"]
    assert set(t) == ({f"Attention_{l}" for l in range(6)} | {f"FeedForward_{l}" for l in range(6)}
                      | {f"PreNorm_{k}" for k in range(12)})
    assert set(t["Attention_0"]["Dense_0"]) == {"kernel"}     # use_bias=False, vit.py:68
    sh = tree_shapes(variables)
    assert sh["pos_embedding"] == (1, 65, 1024) and sh["cls"] == (1, 1, 1024)
    assert sh["Dense_0/kernel"] == (3072, 1024)
    assert sh["Transformer_0/Attention_3/Dense_0/kernel"] == (1024, 3072)
    assert sh["Transformer_0/Attention_3/Dense_1/kernel"] == (1024, 1024)
    assert sh["Transformer_0/FeedForward_5/Dense_0/kernel"] == (1024, 2048)
    assert sh["Dense_1/kernel"] == (1024, 1000)
    assert all(np.asarray(a).dtype == np.float32 for a in flatten_params(variables).values())
    assert v.num_params() == 54_622_184
    # initialisers: zeros / ones / lecun_normal (truncated at 2 sigma, var = 1/fan_in)
    assert not p["pos_embedding"].any() and not p["cls"].any() and not p["Dense_0"]["bias"].any()
    assert (t["PreNorm_0"]["LayerNorm_0"]["scale"] == 1).all()
    k = p["Dense_0"]["kernel"]
    std = np.sqrt(1 / 3072) / 0.87962566103423978
    assert np.abs(k).max() <= 2 * std + 1e-7
    assert abs(k.var() * 3072 - 1.0) < 0.02
    # a PRNGKey-like uint32[2] is accepted as the key
    v2 = v.init({"params": np.array([0, 1], np.uint32)}, img)
    assert tree_shapes(v2) == sh


def test_shape_validation_mirrors_reference_asserts():
    with pytest.raises(AssertionError):                       # vit.py:133-134
        ViT(image_size=250, patch_size=32, num_classes=10, dim=64, depth=1, heads=1, mlp_dim=64).init(
            0, np.zeros((1, 250, 250, 3), np.float32))
    with pytest.raises(AssertionError):                       # vit.py:137
        ViT(pool="max", **TINY).init(0, np.zeros((1, 32, 32, 3), np.float32))
    with pytest.raises(ValueError, match="NCHW"):
        ViT(**TINY).init(0, np.zeros((1, 3, 32, 32), np.float32))


def test_dropout_needs_the_dropout_rng_like_flax():
    v = ViT(dropout=0.1, emb_dropout=0.1, **TINY)             # the demo's config, vit.py:183-184
    variables = v.init({"params": 1, "dropout": 2, "emb_dropout": 3}, np.zeros((1, 32, 32, 3), np.float32))
    with pytest.raises(ValueError, match="rngs=.*dropout"):
        v.apply(variables, np.zeros((1, 32, 32, 3), np.float32))
    with pytest.raises(ValueError, match="rngs=.*dropout"):
        v.apply(variables, np.zeros((1, 32, 32, 3), np.float32), rngs={"emb_dropout": 3})


def test_flatten_accepts_any_mapping_and_leaf_types():
    tree = {"params": MappingProxyType({"a": OrderedDict(kernel=torch.ones(2, 3), bias=[1.0, 2.0, 3.0]),
                                        "cls": np.zeros((1, 1, 4), np.float64)})}
    flat = flatten_params(tree)
    assert set(flat) == {"a/kernel", "a/bias", "cls"}
    for leaf in flat.values():
        a = leaf_to_numpy(leaf)
        assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]


def test_param_specs_cover_every_leaf_once():
    specs = param_specs(**C1)
    paths = ["/".join(p) for p, _, _ in specs]
    assert len(paths) == len(set(paths)) == 4 + 6 * 11 + 4


def test_apply_fails_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    v = ViT(**TINY)
    variables = v.init(0, np.zeros((1, 32, 32, 3), np.float32))
    with pytest.raises((RuntimeError, ImportError)):
        v.apply(variables, np.zeros((1, 32, 32, 3), np.float32))
