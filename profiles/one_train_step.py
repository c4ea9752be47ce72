"""One training step of ViT-B/16 geometry at reduced depth, for `ncu --metrics gpu__time_duration.sum`.
    python profiles/one_train_step.py [depth] [batch] [C2|C4|C5]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
from _util import C2, C4, C5  # noqa: E402
from vit_flax_b200 import init_params  # noqa: E402
from vit_flax_b200.engine import Engine  # noqa: E402

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 2
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
name = sys.argv[3] if len(sys.argv) > 3 else "C2"
cfg = dict(dict(C2=C2, C4=C4, C5=C5)[name], depth=depth)
eng = Engine(precision="fp16", max_batch=batch, **cfg)
eng.load_params(init_params(seed=1, **cfg))
x = torch.randn((batch, cfg["image_size"], cfg["image_size"], 3), device="cuda")
dl = torch.randn((batch, 1000), device="cuda") / batch
for _ in range(2):
    logits = eng.train_forward(x)
    eng.backward(dl)
torch.cuda.synchronize()
print("ok", float(logits.abs().sum()))
