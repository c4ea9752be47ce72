"""Tiny driver for ncu: a few ViT-B/16 batch-256 forwards through the engine (device-resident input).

    python profiles/run_forward.py [steps] [dtype] [batch]
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vit_flax_b200 import init_params, perturb_params  # noqa: E402
from vit_flax_b200.engine import Engine  # noqa: E402

C2 = dict(image_size=224, patch_size=16, num_classes=1000, dim=768, depth=12, heads=12, mlp_dim=3072)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dtype = sys.argv[2] if len(sys.argv) > 2 else "fp16"
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 256
eng = Engine(precision=dtype, max_batch=batch, **C2)
eng.load_params(perturb_params(init_params(seed=1, **C2), seed=2))
x = torch.randn((batch, 224, 224, 3), device="cuda")
for _ in range(steps):
    y = eng.forward(x)
torch.cuda.synchronize()
print("ok", float(y.abs().max()))
