"""Throughput + parity of the other BASELINE configs (they are parity-test cases, not bench lines):
    python profiles/bench_configs.py  ->  markdown rows for profiles/r01_configs.md"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from _util import C2, C3, C4, C5, oracle_logits  # noqa: E402
from vit_flax_b200 import init_params, perturb_params  # noqa: E402
from vit_flax_b200.engine import Engine  # noqa: E402


def flops(cfg):
    P = cfg["patch_size"]; Np = (cfg["image_size"] // P) ** 2; T = Np + 1
    D, L, M, I = cfg["dim"], cfg["depth"], cfg["mlp_dim"], 64 * cfg["heads"]
    return 2 * Np * 3 * P * P * D + L * (2 * T * D * 3 * I + 4 * T * T * I + 2 * T * I * D + 4 * T * D * M) + 2 * D * 1000


torch.set_num_threads(24)
rows = []
for name, cfg, batch in (("C2 ViT-B/16 224", C2, 256), ("C3 ViT-L/16 224", C3, 256), ("C3 ViT-L/16 224", C3, 2048),
                         ("C4 ViT-H/14 224 (inner 1024)", C4, 128), ("C5 ViT-L/16 512", C5, 32), ("C5 ViT-L/16 512", C5, 256)):
    variables = perturb_params(init_params(seed=1, **cfg), seed=2)
    eng = Engine(precision="fp16", max_batch=batch, **cfg)
    eng.load_params(variables)
    s = cfg["image_size"]
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn((batch, s, s, 3), device="cuda", generator=g)
    out = torch.empty((batch, 1000), device="cuda")
    for _ in range(3):
        eng.forward(x, out=out)
    torch.cuda.synchronize()
    n = max(3, int(1.0 / max(1e-3, 0.012 * batch / 256 * flops(cfg) / flops(C2))))
    n = min(n, 20)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        eng.forward(x, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    t0 = time.time()
    want = oracle_logits(variables, x[:2].cpu().numpy(), cfg)
    err = float(np.abs(out[:2].cpu().numpy() - want).max())
    ips = batch / ms * 1e3
    rows.append(f"| {name} | {batch} | {ms:.2f} | {ips:,.0f} | {ips * flops(cfg) / 1e12:.0f} | {err:.1e} | {time.time() - t0:.0f} s |")
    print(rows[-1], flush=True)
    eng.close()
    del x, out
    torch.cuda.empty_cache()
