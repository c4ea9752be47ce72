"""Stand-alone attention driver for timing / ncu.

    python profiles/run_attention.py [batch] [T] [heads] [reps] [dtype]

Prints device time per launch (CUDA events on the launching stream) and the max abs error against
a torch fp32 softmax(QK^T/8)V of the same 16-bit inputs."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vit_flax_b200 import _lib  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 197
heads = int(sys.argv[3]) if len(sys.argv) > 3 else 12
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
dtype = sys.argv[5] if len(sys.argv) > 5 else "fp16"
tdt = torch.float16 if dtype == "fp16" else torch.bfloat16
dt = _lib.DT_F16 if dtype == "fp16" else _lib.DT_BF16
lib = _lib.load()
inner = heads * 64
torch.manual_seed(0)
qkv = (torch.randn((batch * T, 3 * inner), device="cuda") * 1.5).to(tdt)
out = torch.zeros((batch * T, inner), device="cuda", dtype=tdt)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def run():
    _lib.check(lib.vitb200_attention_tc(st, qkv.data_ptr(), out.data_ptr(), batch, T, heads, dt))


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    run()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / reps * 1e3
nb = min(batch, 4)
q, k, v = (t.reshape(nb, T, heads, 64).permute(0, 2, 1, 3).float()
           for t in qkv[: nb * T].split(inner, dim=1))
ref = torch.softmax(q @ k.transpose(-1, -2) * 0.125, dim=-1) @ v
ref = ref.permute(0, 2, 1, 3).reshape(nb * T, inner)
err = (out[: nb * T].float() - ref).abs().max().item()
flops = 4.0 * batch * heads * T * T * 64
print(f"attention batch={batch} T={T} heads={heads} {dtype}: {us:.1f} us  {flops / us / 1e6:.1f} TFLOP/s  "
      f"max abs err {err:.2e}")
