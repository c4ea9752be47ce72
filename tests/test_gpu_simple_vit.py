"""SimpleViT (vit_flax/simple_vit.py) on the same kernels: parity against its numpy restatement."""
import numpy as np
import pytest

from oracle import simple_vit_numpy
from vit_flax_b200 import SimpleViT
from vit_flax_b200.simple_vit import clear_cache, posemb_sincos_2d

pytestmark = pytest.mark.gpu

DEMO = dict(image_size=256, patch_size=32, num_classes=1000, dim=1024, depth=6, heads=16, mlp_dim=2048)  # simple_vit.py:141-150
SMALL = dict(image_size=(32, 64), patch_size=(8, 16), num_classes=10, dim=64, depth=2, heads=2, mlp_dim=128)


def _perturbed(v, img, seed):
    params = v.init({"params": seed}, img)
    rng = np.random.default_rng(seed + 1)

    def walk(node):
        for k, leaf in node.items():
            if isinstance(leaf, dict):
                walk(leaf)
            elif k in ("bias", "scale"):
                node[k] = (leaf + rng.normal(0, 0.05, leaf.shape)).astype(np.float32)
    walk(params["params"])
    return params


@pytest.fixture(autouse=True)
def _fresh():
    yield
    clear_cache()


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("fp16", 2e-2)])
@pytest.mark.parametrize("cfg,batch", [(SMALL, 5), (DEMO, 1)])
def test_simple_vit_matches_oracle(cfg, batch, precision, tol):
    v = SimpleViT(**cfg)
    ih, iw = cfg["image_size"] if isinstance(cfg["image_size"], tuple) else (cfg["image_size"],) * 2
    img = np.random.default_rng(3).standard_normal((batch, 3, ih, iw)).astype(np.float32)   # NCHW
    params = _perturbed(v, img, 4)
    want = simple_vit_numpy.simple_vit_forward(params, img, **cfg)
    got = v.apply(params, img, precision=precision)
    assert got.shape == (batch, cfg["num_classes"]) and got.dtype == np.float32
    assert np.abs(got - want).max() < tol


def test_vit_b16_shape_and_rejections():
    cfg = dict(image_size=224, patch_size=16, num_classes=1000, dim=768, depth=2, heads=12, mlp_dim=3072)
    v = SimpleViT(**cfg)
    img = np.random.default_rng(0).standard_normal((8, 3, 224, 224)).astype(np.float32)
    params = _perturbed(v, img, 1)
    got = v.apply(params, img)
    want = simple_vit_numpy.simple_vit_forward(params, img, **cfg)
    assert np.abs(got - want).max() < 2e-2
    with pytest.raises(ValueError, match="NCHW"):
        v.apply(params, np.zeros((8, 224, 224, 3), np.float32))
    with pytest.raises(NotImplementedError):
        SimpleViT(dim_head=32, **cfg).apply(params, img)


def test_posemb_matches_oracle():
    np.testing.assert_allclose(posemb_sincos_2d(7, 5, 64), simple_vit_numpy.posemb_sincos_2d(7, 5, 64), atol=1e-6)


def test_simple_vit_vjp_matches_finite_differences_of_the_oracle():
    """SimpleViT.vjp (NCHW patchify, no class token, sin/cos table, bias-free LayerNorms, mean pool -- the same
    backward kernels under three config switches).  The numpy oracle has no autograd, so every leaf is checked at
    its largest-gradient entries against central differences of the float64 oracle."""
    import copy
    cfg = SMALL
    v = SimpleViT(**cfg)
    img = np.random.default_rng(5).standard_normal((4, 3, 32, 64)).astype(np.float32)
    params = _perturbed(v, img, 6)
    dl = np.random.default_rng(7).standard_normal((4, cfg["num_classes"]))
    logits, vjp_fn = v.vjp(params, img)
    assert np.abs(logits - simple_vit_numpy.simple_vit_forward(params, img, **cfg)).max() < 2e-2
    grads = vjp_fn(dl.astype(np.float32))

    def leaves(tree, prefix=()):
        for k, x in tree.items():
            if isinstance(x, dict):
                yield from leaves(x, prefix + (k,))
            else:
                yield prefix + (k,), x

    def get(tree, path):
        for k in path:
            tree = tree[k]
        return tree

    f = lambda p: float((simple_vit_numpy.simple_vit_forward(p, img, **cfg) * dl).sum())
    got = dict(leaves(grads["params"]))
    assert set(got) == set(dict(leaves(params["params"])))           # same tree as the parameters
    for path, g in got.items():
        assert g.shape == get(params["params"], path).shape and g.dtype == np.float32
        scale = np.abs(g).max()
        assert scale > 0, path
        for flat_idx in np.argsort(np.abs(g).ravel())[-2:]:
            idx = np.unravel_index(flat_idx, g.shape)
            vals = []
            for sgn in (1, -1):
                p2 = copy.deepcopy(params)
                leaf = get(p2["params"], path[:-1])
                a = np.asarray(leaf[path[-1]], dtype=np.float64).copy()
                a[idx] += sgn * 1e-4
                leaf[path[-1]] = a
                vals.append(f(p2))
            fd = (vals[0] - vals[1]) / 2e-4
            assert abs(g[idx] - fd) < 3e-2 * scale, (path, idx, g[idx], fd)


def test_inline_params_dicts_with_different_weights_are_not_confused():
    """``v.apply({'params': p_i}, x)`` builds a temporary dict per call whose id CPython recycles: the weight cache
    must go by the leaves, not by id(variables).  Also: in-place edits are seen, ``reload`` forces."""
    v = SimpleViT(**SMALL)
    img = np.random.default_rng(0).standard_normal((3, 3, 32, 64)).astype(np.float32)
    p1 = _perturbed(v, img, 10)["params"]
    p2 = _perturbed(v, img, 20)["params"]
    y1 = v.apply({"params": p1}, img, precision="fp32").copy()
    y2 = v.apply({"params": p2}, img, precision="fp32").copy()
    assert np.abs(y1 - simple_vit_numpy.simple_vit_forward({"params": p1}, img, **SMALL)).max() < 1e-4
    assert np.abs(y2 - simple_vit_numpy.simple_vit_forward({"params": p2}, img, **SMALL)).max() < 1e-4
    assert np.abs(y1 - y2).max() > 1e-2
    p2["Dense_0"]["bias"] += 1.0                                   # in-place: every head logit moves by 1
    y3 = v.apply({"params": p2}, img, precision="fp32")
    assert np.abs((y3 - y2) - 1.0).max() < 1e-4
    y4 = v.apply({"params": p2}, img, precision="fp32", reload=True)
    np.testing.assert_array_equal(y3, y4)


def test_adoption_layout_loads_and_gives_the_same_logits_and_gradient_layout():
    v = SimpleViT(**SMALL)
    img = np.random.default_rng(1).standard_normal((2, 3, 32, 64)).astype(np.float32)
    params = _perturbed(v, img, 30)
    p = params["params"]
    adoption = {"params": {"Dense_0": p["Dense_1"], "Transformer_0": p["Transformer_0"],
                           "Sequential_0": {"layers_0": p["LayerNorm_0"], "layers_1": p["Dense_0"]}}}
    ya = v.apply(params, img).copy()
    yb = v.apply(adoption, img)
    np.testing.assert_array_equal(ya, yb)
    dl = np.ones((2, SMALL["num_classes"]), np.float32)
    ga = v.vjp(params, img)[1](dl)["params"]
    gb = v.vjp(adoption, img)[1](dl)["params"]
    assert set(ga) == set(p) and set(gb) == {"Dense_0", "Transformer_0", "Sequential_0"}
    np.testing.assert_array_equal(ga["Dense_1"]["kernel"], gb["Dense_0"]["kernel"])            # patch embedding
    np.testing.assert_array_equal(ga["Dense_0"]["kernel"], gb["Sequential_0"]["layers_1"]["kernel"])   # head
