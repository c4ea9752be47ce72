"""Hang / race hunt for the attention kernel: many launches over varied shapes, each checked against
torch on the first images.  python profiles/stress_attention.py [rounds]"""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vit_flax_b200 import _lib  # noqa: E402

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 3
lib = _lib.load()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
shapes = [(256, 197, 12), (1, 197, 12), (3, 197, 1), (37, 197, 5), (8, 65, 16), (5, 128, 3), (2, 208, 2),
          (9, 129, 4), (4, 1, 2), (64, 17, 7), (300, 197, 12), (2, 64, 1),
          (40, 257, 16), (3, 209, 2), (5, 288, 3), (2, 289, 2), (6, 1025, 4), (1, 2000, 1), (33, 385, 3)]
worst = 0.0
n = 0
for r in range(rounds):
    for (batch, T, heads) in shapes:
        for dt, tdt in ((_lib.DT_F16, torch.float16), (_lib.DT_BF16, torch.bfloat16)):
            inner = heads * 64
            torch.manual_seed(r * 1000 + batch + T)
            qkv = (torch.randn((batch * T, 3 * inner), device="cuda") * 1.5).to(tdt)
            out = torch.full((batch * T, inner), float("nan"), device="cuda", dtype=tdt)
            for _ in range(4):
                _lib.check(lib.vitb200_attention_tc(st, qkv.data_ptr(), out.data_ptr(), batch, T, heads, dt))
                n += 1
            torch.cuda.synchronize()
            err = 0.0
            for b0 in range(0, batch, 32):          # every image, in chunks the reference can hold
                nb = min(32, batch - b0)
                rows = slice(b0 * T, (b0 + nb) * T)
                q, k, v = (t.reshape(nb, T, heads, 64).permute(0, 2, 1, 3).float()
                           for t in qkv[rows].split(inner, dim=1))
                ref = (torch.softmax(q @ k.transpose(-1, -2) * 0.125, dim=-1) @ v).permute(0, 2, 1, 3).reshape(nb * T, inner)
                e = (out[rows].float() - ref).abs().max().item()
                if e > err:
                    err, where = e, b0
            
            tail_ok = bool(torch.isfinite(out.float()).all())
            tol = 4e-3 if tdt == torch.float16 else 3e-2
            assert err < tol and tail_ok, (batch, T, heads, tdt, err, tail_ok, where)
            worst = max(worst, err / tol)
print(f"stress ok: {n} launches, worst err/tol {worst:.2f}")
