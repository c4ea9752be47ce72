"""Isolated timings of the backward kernels that have per-kernel C entry points, at ViT-B/16 batch-256 sizes.
    python profiles/run_bwd_kernels.py"""
import ctypes as C
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vit_flax_b200 import _lib  # noqa: E402

lib = _lib.load()
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
R, D, H, heads, T, B = 50432, 768, 3072, 12, 197, 256
dt, tdt = _lib.DT_F16, torch.float16


def timeit(name, fn, bytes_=None, flops=None, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) / n * 1e3
    extra = (f"  {bytes_ / us / 1e6:.2f} TB/s" if bytes_ else "") + (f"  {flops / us / 1e6:.0f} TFLOP/s" if flops else "")
    print(f"{name:44s} {us:8.1f} us{extra}")


x = torch.randn((R, D), device="cuda")
dy = torch.randn((R, D), device="cuda").to(tdt)
g = torch.ones(D, device="cuda")
dx = torch.zeros((R, D), device="cuda")
dg, db = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
timeit("layernorm_bwd [50432, 768] accumulate",
       lambda: _lib.check(lib.vitb200_layernorm_bwd(st(), dy.data_ptr(), x.data_ptr(), g.data_ptr(), dx.data_ptr(), dg.data_ptr(),
                                                    db.data_ptr(), R, D, dt, 1e-6, 1)), bytes_=R * D * (2 + 4 + 4 + 4))
qkv = torch.randn((R, 3 * heads * 64), device="cuda").to(tdt)
do = torch.randn((R, heads * 64), device="cuda").to(tdt)
dqkv = torch.empty_like(qkv)
o = torch.randn((R, heads * 64), device="cuda").to(tdt)
timeit("attention_bwd batch 256, T 197, 12 heads",
       lambda: _lib.check(lib.vitb200_attention_bwd(st(), qkv.data_ptr(), o.data_ptr(), do.data_ptr(), dqkv.data_ptr(), B, T, heads, dt)),
       flops=5 * 2 * T * T * 64 * B * heads)   # per-kernel entry point: includes the forward kernel re-run for the log-sum-exp
for (M, N, name) in ((D, 3 * D, "to_qkv"), (D, D, "to_out"), (D, H, "ff1"), (H, D, "ff2")):
    X = torch.randn((R, M), device="cuda").to(tdt)
    dY = torch.randn((R, N), device="cuda").to(tdt)
    dW = torch.zeros((M, N), device="cuda")
    tiles = ((M + 255) // 256) * ((N + 255) // 256)
    sp = max(1, 148 // tiles)
    timeit(f"wgrad {name} [{M} x {N}], K = {R}, split {sp}",
           lambda: _lib.check(lib.vitb200_gemm_tc_wgrad(st(), X.data_ptr(), dY.data_ptr(), dW.data_ptr(), M, N, R, sp, dt)),
           flops=2 * M * N * R)
    del X, dY, dW
