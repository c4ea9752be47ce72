"""Small-batch latency with and without graph replay of the forward (VITB200_GRAPH=0|unset).
    VITB200_GRAPH=0 python profiles/latency_graph.py; python profiles/latency_graph.py"""
import os
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
from _util import C1, C2  # noqa: E402
from vit_flax_b200 import init_params, perturb_params  # noqa: E402
from vit_flax_b200.engine import Engine  # noqa: E402

print("VITB200_GRAPH =", os.environ.get("VITB200_GRAPH", "(auto)"))
for name, cfg in (("C1 README", C1), ("C2 ViT-B/16", C2)):
    for batch in (1, 8, 32):
        eng = Engine(precision="fp16", max_batch=batch, **cfg)
        eng.load_params(perturb_params(init_params(seed=1, **cfg), seed=2))
        s = cfg["image_size"]
        x = torch.randn((batch, s, s, 3), device="cuda")
        out = torch.empty((batch, 1000), device="cuda")
        for _ in range(20):
            eng.forward(x, out=out)
        torch.cuda.synchronize()
        n = 200
        t0 = time.perf_counter()
        for _ in range(n):
            eng.forward(x, out=out)
            torch.cuda.synchronize()                      # one request at a time: latency, not throughput
        lat = (time.perf_counter() - t0) / n * 1e6
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            eng.forward(x, out=out)
        e1.record()
        torch.cuda.synchronize()
        print(f"{name} batch {batch}: {lat:.0f} us per synchronous forward, {e0.elapsed_time(e1) / n * 1e3:.0f} us back to back, "
              f"checksum {float(out.double().abs().sum()):.6f}")
        eng.close()
