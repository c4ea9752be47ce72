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
