"""Debug timeline of the tcgen05 attention adjoint (CTA 0): build with VITB200_TRACE=1 python -m vit_flax_b200.build --force

    python profiles/trace_attention_bwd.py [batch] [T] [heads]
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
qkv = (torch.randn((batch * T, 3 * inner), device="cuda") * 1.2).to(torch.float16)
d_out = torch.randn((batch * T, inner), device="cuda").to(torch.float16)
out = torch.zeros((batch * T, inner), device="cuda", dtype=torch.float16)
dqkv = torch.zeros((batch * T, 3 * inner), device="cuda", dtype=torch.float16)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
_lib.check(lib.vitb200_attention_tc(st, qkv.data_ptr(), out.data_ptr(), batch, T, heads, _lib.DT_F16))
for _ in range(2):
    _lib.check(lib.vitb200_attention_bwd(st, qkv.data_ptr(), out.data_ptr(), d_out.data_ptr(), dqkv.data_ptr(), batch, T, heads, _lib.DT_F16))
torch.cuda.synchronize()
tr = np.zeros((16, 64), np.int64)
lib.vitb200_debug_attention_bwd_trace.argtypes = [C.c_void_p, C.c_int]
assert lib.vitb200_debug_attention_bwd_trace(tr.ctypes.data, tr.size) == 0
names = ["iss_wait", "pds_seen", "dV_done", "half1_iss", "blk_done", "m_h0_rdy", "m_h0_calc", "m_tiles", "m_h1_rdy", "m_publish",
         "load_iss", "half0_iss", "m_h0_ld", "m_h1_ld", "m_h1_calc", "m_stored"]
t0 = tr[10, 0]
print("blk  " + " ".join(f"{n:>9s}" for n in names))
for i in range(0, 40):
    print(f"{i:4d} " + " ".join(f"{int(tr[e, i] - t0):9d}" for e in range(16)))
print("per-block period (cycles):", (tr[9, 36] - tr[9, 12]) / 24.0)
