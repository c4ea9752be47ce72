"""Small end-to-end run for compute-sanitizer: tiny ViT (all three precisions) + one ViT-B/16-shaped
layer at batch 2, one T=257 and one T=1025 attention call.  Checks results like smoke()."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from _util import C2, C4, TINY, images_for, load_golden, oracle_logits  # noqa: E402
from vit_flax_b200 import ViT, init_params, perturb_params  # noqa: E402

variables, meta = load_golden("tiny_cls.npz")
for precision, tol in (("fp32", 1e-4), ("fp16", 2e-2), ("bf16", 5e-2)):
    y = ViT(**TINY).apply(variables, meta["images"], precision=precision)
    assert np.abs(y - meta["logits"]).max() < tol, precision
for cfg, batch in ((dict(C2, depth=1), 2), (dict(C4, depth=1), 1)):
    v = perturb_params(init_params(seed=1, **cfg), seed=2)
    img = images_for(cfg, batch)
    want = oracle_logits(v, img, cfg)
    got = ViT(**cfg).apply(v, img)
    assert np.abs(got - want).max() < 2e-2
torch.cuda.synchronize()
print("sanitize_small ok")
