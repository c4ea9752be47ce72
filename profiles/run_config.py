"""Per-kernel breakdown of one forward for any BASELINE config.
    python profiles/run_config.py C2|C3|C4|C5 batch [dtype]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
from _util import C2, C3, C4, C5  # noqa: E402
from vit_flax_b200 import init_params, perturb_params  # noqa: E402
from vit_flax_b200.engine import Engine  # noqa: E402

name = sys.argv[1]
batch = int(sys.argv[2])
dtype = sys.argv[3] if len(sys.argv) > 3 else "fp16"
cfg = dict(C2=C2, C3=C3, C4=C4, C5=C5)[name]
eng = Engine(precision=dtype, max_batch=batch, **cfg)
eng.load_params(perturb_params(init_params(seed=1, **cfg), seed=2))
s = cfg["image_size"]
x = torch.randn((batch, s, s, 3), device="cuda")
out = torch.empty((batch, 1000), device="cuda")
for _ in range(3):
    eng.forward(x, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 5
for _ in range(n):
    eng.forward(x, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
prof = eng.profile_forward(x, out=out)
tot = sum(m for m, _ in prof.values())
print(f"{name} batch {batch} {dtype}: {ms:.2f} ms/forward = {batch / ms * 1e3:.0f} img/s")
print("  " + "  ".join(f"{k} {m:.2f} ({100 * m / tot:.0f}%)" for k, (m, c) in prof.items() if m > 0.005 * tot))
