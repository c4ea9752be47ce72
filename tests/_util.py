"""Shared helpers for the test-suite (CPU oracle = the checker, never the product path)."""
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"

TINY = dict(image_size=32, patch_size=8, num_classes=8, dim=64, depth=2, heads=2, mlp_dim=128)
TINY_MEAN = dict(image_size=(16, 32), patch_size=(8, 16), num_classes=16, dim=64, depth=1, heads=1, mlp_dim=64)
C1 = dict(image_size=256, patch_size=32, num_classes=1000, dim=1024, depth=6, heads=16, mlp_dim=2048)
C2 = dict(image_size=224, patch_size=16, num_classes=1000, dim=768, depth=12, heads=12, mlp_dim=3072)
C3 = dict(image_size=224, patch_size=16, num_classes=1000, dim=1024, depth=24, heads=16, mlp_dim=4096)
C4 = dict(image_size=224, patch_size=14, num_classes=1000, dim=1280, depth=32, heads=16, mlp_dim=5120)
C5 = dict(image_size=512, patch_size=16, num_classes=1000, dim=1024, depth=24, heads=16, mlp_dim=4096)


def load_golden(name):
    """-> (variables or None, dict of the non-param arrays)."""
    z = np.load(GOLDEN / name)
    meta, tree = {}, {}
    for k in z.files:
        if k in ("images", "logits", "tokens", "init_seed", "perturb_seed", "image_seed"):
            meta[k] = z[k]
            continue
        node = tree
        parts = k.split(".")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = z[k]
    return ({"params": tree} if tree else None), meta


def images_for(cfg, batch, seed=0, channels=3):
    ih, iw = cfg["image_size"] if isinstance(cfg["image_size"], tuple) else (cfg["image_size"],) * 2
    return np.random.default_rng(seed).standard_normal((batch, ih, iw, channels)).astype(np.float32)


def oracle_logits(variables, images, cfg, **kw):
    """The checker (torch-CPU restatement of vit.py) for scripts outside tests/ that want a parity figure next to
    their timings: they import THIS, so that everything under oracle/ stays reachable from tests/, smoke() and
    bench.py's CPU legs only."""
    from oracle import vit_torch
    return vit_torch.vit_forward(vit_torch.tree_to_torch(variables), images, **kw, **cfg).numpy()
