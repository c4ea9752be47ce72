"""Stand-alone tcgen05 GEMM driver for timing / ncu.

    python profiles/run_gemm.py M N K epilogue cta_group [reps]

Prints TFLOP/s from CUDA events (device time on the launching stream)."""
import ctypes as C
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vit_flax_b200 import _lib  # noqa: E402

M, N, K, epi, cg = (int(a) for a in sys.argv[1:6])
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 20
os.environ["VITB200_GEMM_CTA_GROUP"] = str(cg)
lib = _lib.load()
A = torch.randn((M, K), device="cuda").to(torch.float16)
Wt = (torch.randn((N, K), device="cuda") / K ** 0.5).to(torch.float16)
bias = torch.randn(N, device="cuda")
out16 = epi in (0, 1)
Cb = torch.zeros((M, N), device="cuda", dtype=torch.float16 if out16 else torch.float32)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def run():
    _lib.check(lib.vitb200_gemm_tc(st, A.data_ptr(), Wt.data_ptr(), bias.data_ptr(), Cb.data_ptr(),
                                   M, N, K, epi, None, 0, _lib.DT_F16))


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"M={M} N={N} K={K} epi={epi} cta_group={cg}: {ms * 1e3:.1f} us  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s")
