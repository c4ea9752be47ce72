"""Debug timeline of the attention kernel (CTA 0): build with VITB200_TRACE=1 python -m vit_flax_b200.build --force

    python profiles/trace_attention.py [batch] [T] [heads]
"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vit_flax_b200 import _lib  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 197
heads = int(sys.argv[3]) if len(sys.argv) > 3 else 12
lib = _lib.load()
inner = heads * 64
qkv = (torch.randn((batch * T, 3 * inner), device="cuda") * 1.5).to(torch.float16)
out = torch.zeros((batch * T, inner), device="cuda", dtype=torch.float16)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(1):
    _lib.check(lib.vitb200_attention_tc(st, qkv.data_ptr(), out.data_ptr(), batch, T, heads, _lib.DT_F16))
torch.cuda.synchronize()
tr = np.zeros((12, 64), np.int64)
lib.vitb200_debug_attention_trace.argtypes = [C.c_void_p, C.c_int]
assert lib.vitb200_debug_attention_trace(tr.ctypes.data, tr.size) == 0
names = ["qk_load", "v_load", "S_issue", "PV_issue", "sm_begin", "s_ready", "pass1_end", "p_ready", "epi_begin",
         "p2_begin", "stored", "p2_end"]
t0 = tr[0, 0]
print("item " + " ".join(f"{n:>9s}" for n in names))
for i in range(0, 30):
    print(f"{i:4d} " + " ".join(f"{int(tr[e, i] - t0):9d}" for e in range(12)))
print("per-item period (cycles):", (tr[7, 30] - tr[7, 10]) / 20.0)
d = lambda a, b: float(np.mean([tr[b, i] - tr[a, i] for i in range(8, 32)]))
print(f"means over items 8..31: S_issue->s_ready {d(2, 5):.0f}  s_ready->pass1_end {d(5, 6):.0f}  pass1_end->p2_begin {d(6, 9):.0f}  "
      f"p2_begin->p2_end {d(9, 11):.0f}  p2_end->p_ready {d(11, 7):.0f}  PV_issue->p_ready {d(3, 7):.0f}  p_ready->stored {d(7, 10):.0f}")
