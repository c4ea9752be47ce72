"""Regenerate the committed golden vectors:  python tests/golden/make_golden.py

These vectors come from the float64 oracle (oracle/vit_numpy.py), cross-checked
against the independent torch restatement before being written: they pin the
ORACLE (and the param initialiser) against drift.  The outputs of the reference
itself on the same inputs are in ref_vit.npz (make_reference_golden.py), and
tests/test_reference_run.py checks that the two agree.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import vit_numpy, vit_torch  # noqa: E402
from vit_flax_b200.params import flatten_params, init_params, perturb_params  # noqa: E402

HERE = Path(__file__).resolve().parent

TINY = dict(image_size=32, patch_size=8, num_classes=8, dim=64, depth=2, heads=2, mlp_dim=128)
TINY_MEAN = dict(image_size=(16, 32), patch_size=(8, 16), num_classes=16, dim=64, depth=1, heads=1, mlp_dim=64)
C1 = dict(image_size=256, patch_size=32, num_classes=1000, dim=1024, depth=6, heads=16, mlp_dim=2048)


def check(cfg, variables, images, y, pool="cls"):
    yt = vit_torch.vit_forward(vit_torch.tree_to_torch(variables), images, pool=pool, **cfg).numpy()
    err = np.abs(yt - y).max()
    assert err < 2e-5, err
    return err


def main():
    # 1. tiny config, params stored (pins the oracle independently of the initialiser)
    v = perturb_params(init_params(seed=11, **TINY), seed=12)
    img = np.random.default_rng(13).standard_normal((3, 32, 32, 3)).astype(np.float32)
    y, tok = vit_numpy.vit_forward(v, img, return_tokens=True, **TINY)
    print("tiny torch-vs-numpy", check(TINY, v, img, y))
    flat = {k.replace("/", "."): a for k, a in flatten_params(v).items()}
    np.savez_compressed(HERE / "tiny_cls.npz", images=img, logits=y, tokens=tok, **flat)

    # 2. non-square image / patch, mean pool, heads == 1 and dim == 64 => no to_out (vit.py:65)
    v = perturb_params(init_params(seed=21, **TINY_MEAN), seed=22)
    img = np.random.default_rng(23).standard_normal((2, 16, 32, 3)).astype(np.float32)
    y = vit_numpy.vit_forward(v, img, pool="mean", **TINY_MEAN)
    print("tiny-mean torch-vs-numpy", check(TINY_MEAN, v, img, y, pool="mean"))
    flat = {k.replace("/", "."): a for k, a in flatten_params(v).items()}
    np.savez_compressed(HERE / "tiny_mean.npz", images=img, logits=y, **flat)

    # 3. README config (C1), params regenerated from seeds (too large to store)
    v = perturb_params(init_params(seed=1, **C1), seed=2)
    img = np.random.default_rng(0).standard_normal((1, 256, 256, 3)).astype(np.float32)
    y = vit_numpy.vit_forward(v, img, **C1)
    print("C1 torch-vs-numpy", check(C1, v, img, y))
    np.savez_compressed(HERE / "c1_logits.npz", logits=y, init_seed=1, perturb_seed=2, image_seed=0)


if __name__ == "__main__":
    main()
