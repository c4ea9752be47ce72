"""Race hunt for the backward pass: many training steps on two alternating batches; the gradients of a batch must come
out the same every time (up to the order of fp32 atomics / split-K reduce-adds), at full size and at ragged sizes.
    python profiles/stress_backward.py [steps]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
from _util import C2  # noqa: E402
from vit_flax_b200 import init_params, perturb_params  # noqa: E402
from vit_flax_b200.engine import Engine  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
for cfg, batch in ((dict(C2, depth=4), 256), (dict(C2, depth=2, image_size=240), 37)):     # T = 197 resident, T = 226 streamed
    eng = Engine(precision="fp16", max_batch=batch, **cfg)
    eng.load_params(perturb_params(init_params(seed=1, **cfg), seed=2))
    s = cfg["image_size"]
    g = torch.Generator(device="cuda").manual_seed(0)
    data = [(torch.randn((batch, s, s, 3), device="cuda", generator=g), torch.randn((batch, 1000), device="cuda", generator=g) / batch)
            for _ in range(2)]
    ref = [None, None]
    worst = 0.0
    for i in range(steps):
        x, dl = data[i & 1]
        eng.train_forward(x)
        eng.backward(dl)
        flat = eng.grads_flat()
        assert torch.isfinite(flat).all(), f"non-finite gradient at step {i}"
        if ref[i & 1] is None:
            ref[i & 1] = flat.clone()
        else:
            err = float((flat - ref[i & 1]).abs().max() / ref[i & 1].abs().max())
            worst = max(worst, err)
            assert err < 1e-3, f"step {i}: gradients of the same batch moved by {err:.2e}"
    print(f"T = {eng.tokens}, batch {batch}, depth {cfg['depth']}: {steps} steps, same-batch gradients repeat to {worst:.1e} (max rel)")
    eng.close()
