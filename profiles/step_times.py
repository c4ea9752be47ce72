"""Per-step device times of the ViT-B/16 batch-256 forward (one CUDA event pair per step) + clocks.
    python profiles/step_times.py [steps] [dtype]"""
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vit_flax_b200 import init_params, perturb_params  # noqa: E402
from vit_flax_b200.engine import Engine  # noqa: E402

C2 = dict(image_size=224, patch_size=16, num_classes=1000, dim=768, depth=12, heads=12, mlp_dim=3072)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
dtype = sys.argv[2] if len(sys.argv) > 2 else "fp16"
eng = Engine(precision=dtype, max_batch=256, **C2)
eng.load_params(perturb_params(init_params(seed=1, **C2), seed=2))
x = torch.randn((256, 224, 224, 3), device="cuda")
out = torch.empty((256, 1000), device="cuda")
rows = []
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.active",
                      "--format=csv,noheader,nounits", "-lms", "20", "-i", "0"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [rows.append(l.strip()) for l in p.stdout], daemon=True).start()
for _ in range(5):
    eng.forward(x, out=out)
torch.cuda.synchronize()
time.sleep(0.5)
n0 = len(rows)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
ev[0].record()
for i in range(steps):
    eng.forward(x, out=out)
    ev[i + 1].record()
torch.cuda.synchronize()
n1 = len(rows)
p.terminate()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
print("first 10 steps:", [round(m, 2) for m in ms[:10]])
print("last 10 steps :", [round(m, 2) for m in ms[-10:]])
print(f"mean first 20 {sum(ms[:20]) / 20:.3f} ms, mean last 20 {sum(ms[-20:]) / 20:.3f} ms, min {min(ms):.3f}")
print("clock samples during run:", rows[n0:n1][:: max(1, (n1 - n0) // 12)])
