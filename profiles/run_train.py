"""Training step (train_forward + backward) of a BASELINE config: device time per phase.
    python profiles/run_train.py C2 256 [fp16|bf16] [steps]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
from _util import C2, C3, C4, C5  # noqa: E402
from vit_flax_b200 import init_params, perturb_params  # noqa: E402
from vit_flax_b200.engine import Engine  # noqa: E402

name, batch = sys.argv[1], int(sys.argv[2])
dtype = sys.argv[3] if len(sys.argv) > 3 else "fp16"
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
cfg = dict(C2=C2, C3=C3, C4=C4, C5=C5)[name]
eng = Engine(precision=dtype, max_batch=batch, **cfg)
eng.load_params(perturb_params(init_params(seed=1, **cfg), seed=2))
s = cfg["image_size"]
x = torch.randn((batch, s, s, 3), device="cuda")
logits = torch.empty((batch, 1000), device="cuda")
dl = torch.randn((batch, 1000), device="cuda") / batch
for _ in range(2):
    eng.train_forward(x, out=logits)
    eng.backward(dl)
torch.cuda.synchronize()
print(f"memory in use after warm-up: {torch.cuda.mem_get_info()[1] / 2**30 - torch.cuda.mem_get_info()[0] / 2**30:.1f} GiB")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * steps + 1)]
ev[0].record()
for i in range(steps):
    eng.train_forward(x, out=logits)
    ev[2 * i + 1].record()
    eng.backward(dl)
    ev[2 * i + 2].record()
torch.cuda.synchronize()
fwd = sum(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(steps)) / steps
bwd = sum(ev[2 * i + 1].elapsed_time(ev[2 * i + 2]) for i in range(steps)) / steps
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    eng.forward(x, out=logits)
e1.record()
torch.cuda.synchronize()
inf = e0.elapsed_time(e1) / steps
# 3 x forward FLOPs (dgrad + wgrad per Dense, 2.5x attention) is the usual training-step estimate
print(f"{name} batch {batch} {dtype}: train_forward {fwd:.2f} ms, backward {bwd:.2f} ms, step {fwd + bwd:.2f} ms "
      f"= {batch / (fwd + bwd) * 1e3:.0f} img/s;  inference forward {inf:.2f} ms;  backward / forward = {bwd / inf:.2f}")
