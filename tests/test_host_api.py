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


# ---- weight-cache identity (runtime.ParamsStamp) and the SimpleViT params layouts: host logic, no GPU ----
def test_params_stamp_sees_recycled_ids_and_inplace_edits():
    from vit_flax_b200.params import init_params
    from vit_flax_b200.runtime import ParamsStamp
    p1 = init_params(seed=1, **TINY)["params"]
    p2 = init_params(seed=2, **TINY)["params"]
    # the usual flax call builds a temporary {'params': p} per call: CPython may hand both the same id
    s1 = ParamsStamp({"params": p1})
    s2 = ParamsStamp({"params": p2})
    assert not s2.matches(s1) and s1.matches(ParamsStamp({"params": p1}))
    assert not s1.matches(None)
    # in-place edits of a host leaf are seen by the content probe, a torch leaf by its version counter
    before = ParamsStamp(p1)
    p1["Dense_0"]["kernel"] *= 0.5
    assert not ParamsStamp(p1).matches(before)
    t = {k: torch.as_tensor(np.array(v)) if not isinstance(v, dict) else v for k, v in p2.items()}
    before = ParamsStamp(t)
    assert ParamsStamp(t).matches(before)
    t["cls"].add_(1.0)
    assert not ParamsStamp(t).matches(before)
    # a different tree structure never matches
    assert not ParamsStamp({"a": np.zeros(3, np.float32)}).matches(ParamsStamp({"b": np.zeros(3, np.float32)}))


def test_simple_vit_init_layout_is_flax_construction_scope_and_both_layouts_load():
    """simple_vit.py:117-126: LayerNorm and head Dense are constructed (inside nn.Sequential([...])) in SimpleViT's
    scope BEFORE the patch Dense => LayerNorm_0, Dense_0 = head, Dense_1 = patch embedding, no Sequential_0 params.
    The adoption layout (Dense_0 = patch, Sequential_0/layers_*) is accepted too and maps to the same weights."""
    from oracle import simple_vit_numpy
    from vit_flax_b200 import SimpleViT
    cfg = dict(image_size=(32, 64), patch_size=(8, 16), num_classes=10, dim=64, depth=2, heads=2, mlp_dim=128)
    v = SimpleViT(**cfg)
    img = np.random.default_rng(0).standard_normal((2, 3, 32, 64)).astype(np.float32)
    variables = v.init({"params": 3}, img)
    p = variables["params"]
    assert set(p) == {"LayerNorm_0", "Dense_0", "Dense_1", "Transformer_0"}
    assert p["Dense_0"]["kernel"].shape == (64, 10) and p["Dense_1"]["kernel"].shape == (8 * 16 * 3, 64)
    assert set(p["LayerNorm_0"]) == {"scale"}                                        # use_bias=False
    assert set(p["Transformer_0"]) == {f"{m}_{l}" for m in ("Attention", "FeedForward") for l in range(2)}
    assert set(p["Transformer_0"]["Attention_0"]) == {"LayerNorm_0", "Dense_0", "Dense_1"}
    assert set(p["Transformer_0"]["Attention_0"]["Dense_1"]) == {"kernel"}           # to_out has no bias
    adoption = {"params": {"Dense_0": p["Dense_1"], "Transformer_0": p["Transformer_0"],
                           "Sequential_0": {"layers_0": p["LayerNorm_0"], "layers_1": p["Dense_0"]}}}
    assert v._split_tree(variables)[0] == "construction" and v._split_tree(adoption)[0] == "adoption"
    fa, fb = flatten_params(v._engine_tree(variables)), flatten_params(v._engine_tree(adoption))
    assert sorted(fa) == sorted(fb) and all(np.array_equal(fa[k], fb[k]) for k in fa)
    np.testing.assert_array_equal(simple_vit_numpy.simple_vit_forward(variables, img, **cfg),
                                  simple_vit_numpy.simple_vit_forward(adoption, img, **cfg))
    # a tree whose head / patch kernels are swapped (the other reading of the names) fails loudly
    swapped = {"params": dict(p, Dense_0=p["Dense_1"], Dense_1=p["Dense_0"])}
    with pytest.raises(ValueError, match="kernel shapes"):
        v._engine_tree(swapped)
    with pytest.raises(ValueError, match="lacks"):
        v._engine_tree({"params": {"Transformer_0": p["Transformer_0"]}})
    # gradients come back in the caller's layout
    flat = {k: np.zeros(s, np.float32) for k, s in tree_shapes(v._engine_tree(variables)).items()}
    assert set(v._grads_tree(flat)["params"]) == set(p)
    assert set(v._grads_tree(flat, "adoption")["params"]) == {"Dense_0", "Transformer_0", "Sequential_0"}
