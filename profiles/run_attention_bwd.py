"""Attention adjoint alone (per-kernel C entry point), for ncu / timing:  python profiles/run_attention_bwd.py B T heads reps [fp16|bf16]
The entry point has no forward log-sum-exp, so every call runs the small statistics kernel first and then
attention_bwd_tc5_kernel (T <= 208) -- profile with -k regex:attention_bwd_tc5."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vit_flax_b200 import _lib  # noqa: E402

B, T, heads, reps = (int(a) for a in sys.argv[1:5])
fmt = sys.argv[5] if len(sys.argv) > 5 else "fp16"
dt, tdt = (_lib.DT_F16, torch.float16) if fmt == "fp16" else (_lib.DT_BF16, torch.bfloat16)
lib = _lib.load()
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
inner = heads * 64
qkv = torch.randn((B * T, 3 * inner), device="cuda").to(tdt)
do = torch.randn((B * T, inner), device="cuda").to(tdt)
o = torch.empty((B * T, inner), device="cuda", dtype=tdt)
dqkv = torch.empty_like(qkv)
_lib.check(lib.vitb200_attention_tc(st(), qkv.data_ptr(), o.data_ptr(), B, T, heads, dt))
for _ in range(2):
    _lib.check(lib.vitb200_attention_bwd(st(), qkv.data_ptr(), o.data_ptr(), do.data_ptr(), dqkv.data_ptr(), B, T, heads, dt))
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    _lib.check(lib.vitb200_attention_bwd(st(), qkv.data_ptr(), o.data_ptr(), do.data_ptr(), dqkv.data_ptr(), B, T, heads, dt))
b.record()
torch.cuda.synchronize()
us = a.elapsed_time(b) / reps * 1e3
print(f"attention_bwd (stats + tcgen05 kernel) B={B} T={T} heads={heads} {fmt}: {us:.1f} us per call, "
      f"{5 * 2 * T * T * 64 * B * heads / us / 1e6:.0f} TFLOP/s of useful work")
