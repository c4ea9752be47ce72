"""Sustained (power-capped) throughput of the tcgen05 GEMM next to cuBLAS (torch.matmul) on the same
shape: each runs back to back for `secs` seconds; reports TFLOP/s over the last half.
    python profiles/sustained_gemm.py M N K epilogue [secs]"""
import ctypes as C
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vit_flax_b200 import _lib  # noqa: E402

M, N, K, epi = (int(a) for a in sys.argv[1:5])
secs = float(sys.argv[5]) if len(sys.argv) > 5 else 3.0
lib = _lib.load()
A = torch.randn((M, K), device="cuda").to(torch.float16)
Wt = (torch.randn((N, K), device="cuda") / K ** 0.5).to(torch.float16)
bias = torch.randn(N, device="cuda")
out16 = epi in (0, 1)
Cb = torch.zeros((M, N), device="cuda", dtype=torch.float16 if out16 else torch.float32)
Cc = torch.empty((M, N), device="cuda", dtype=torch.float16)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ours():
    _lib.check(lib.vitb200_gemm_tc(st, A.data_ptr(), Wt.data_ptr(), bias.data_ptr(), Cb.data_ptr(),
                                   M, N, K, epi, None, 0, _lib.DT_F16))


def cublas():
    torch.matmul(A, Wt.t(), out=Cc)


def sustained(fn, label):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t_end = time.time() + secs
    rates = []
    while time.time() < t_end:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            fn()
        e1.record()
        torch.cuda.synchronize()
        rates.append(2.0 * M * N * K * 50 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    tail = rates[len(rates) // 2:]
    print(f"{label:8s} M={M} N={N} K={K}: first {rates[0]:.0f}  sustained {sum(tail) / len(tail):.0f} TFLOP/s ({len(rates)} batches)")


sustained(cublas, "cuBLAS")
time.sleep(2)
sustained(ours, f"ours e{epi}")
