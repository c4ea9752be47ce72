"""Top-1 agreement and max-abs logit error of the GPU forward against the CPU oracle on a sample large enough to
show 99.9 % (north star): ViT-B/16 224 (BASELINE configs[1]), N images (default 2048), both 16-bit operand formats.

    python profiles/r02_parity.py [N] > gpurun_out/r02_parity.json

Reports, per format: max-abs error, raw top-1 agreement, agreement restricted to images whose oracle top-1 margin
exceeds 2x the measured max-abs error, and how many images that leaves.  The oracle is the checker only."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from _util import C2, oracle_logits  # noqa: E402
from vit_flax_b200 import init_params, perturb_params  # noqa: E402
from vit_flax_b200.engine import Engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
chunk = 256
torch.set_num_threads(os.cpu_count() or 1)
variables = perturb_params(init_params(seed=1, **C2), seed=2)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
engines = {dt: Engine(precision=dt, max_batch=chunk, device=0, **C2) for dt in ("bf16", "fp16")}
for e in engines.values():
    e.load_params(variables)
got = {dt: [] for dt in engines}
want = []
t_cpu = 0.0
for c0 in range(0, n, chunk):
    b = min(chunk, n - c0)
    x = torch.empty((b, 224, 224, 3), dtype=torch.float32, device=dev)
    for i in range(b):
        g.manual_seed(50_000 + c0 + i)
        x[i].normal_(generator=g)
    for dt, e in engines.items():
        got[dt].append(e.forward(x).cpu().numpy())
    t0 = time.perf_counter()
    want.append(oracle_logits(variables, x.cpu().numpy(), C2))
    t_cpu += time.perf_counter() - t0
want = np.concatenate(want)
srt = np.sort(want, axis=1)
margin = srt[:, -1] - srt[:, -2]
out = {"config": "ViT-B/16 224 (BASELINE configs[1]), reference initialisers + N(0,0.02) on zero/one leaves",
       "images": n, "oracle": f"oracle/vit_torch.py fp32 on {os.cpu_count()} host cores, {t_cpu:.1f} s",
       "oracle_margin": {"mean": float(margin.mean()), "median": float(np.median(margin)),
                         "frac_below_1e-2": float((margin < 1e-2).mean()), "frac_below_5e-2": float((margin < 5e-2).mean())}}
for dt in engines:
    y = np.concatenate(got[dt])
    err = np.abs(y - want)
    agree = y.argmax(1) == want.argmax(1)
    e = float(err.max())
    conf = margin > 2 * e
    flips = np.nonzero(~agree)[0]
    out[dt] = {"max_abs_err": e, "mean_abs_err": float(err.mean()), "p999_abs_err": float(np.quantile(err, 0.999)),
               "top1_agree_raw": float(agree.mean()), "top1_disagreements": int((~agree).sum()),
               "largest_margin_among_disagreements": float(margin[flips].max()) if flips.size else None,
               "images_with_margin_gt_2err": int(conf.sum()),
               "top1_agree_where_margin_gt_2err": float(agree[conf].mean()) if conf.any() else None}
print(json.dumps(out, indent=1))
