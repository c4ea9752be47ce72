"""Per-kernel parity on a real B200, called through the C ABI (ctypes), checked against the
CPU oracle (numpy float64 restatement of vit.py).  Tolerances are stated per test:
bf16-operand kernels are compared on bf16-rounded inputs, so the only error left is fp32
accumulation order plus the final bf16 rounding where the output is bf16."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import vit_numpy
from vit_flax_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib(lib_built):
    assert lib_built.vitb200_device_count() >= 1, "no sm_100 device"
    return lib_built


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dev(a, dtype=torch.float32):
    return torch.as_tensor(np.asarray(a)).to("cuda", dtype).contiguous()


def bf16_round(a):
    return torch.as_tensor(np.asarray(a, np.float32)).to(torch.bfloat16).float().numpy().astype(np.float64)


GEMM_SHAPES = [
    (128, 256, 64),       # exactly one tile, one k-block
    (128, 256, 512),      # pipeline wraps the 4-stage ring twice
    (300, 264, 136),      # ragged M, N, K tails (TMA zero fill + masked stores)
    (1576, 2304, 768),    # ViT-B to_qkv at batch 8
    (1576, 768, 3072),    # ViT-B FF down-projection at batch 8 (48 k-blocks)
    (40000, 768, 768),    # > 148*2 tiles: persistent loop + TMEM double buffering, many waves
]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("epi", [_lib.EPI_STORE_BF16, _lib.EPI_BIAS_GELU_BF16,
                                 _lib.EPI_BIAS_RESID_F32, _lib.EPI_BIAS_F32])
def test_gemm_bf16_epilogues(lib, M, N, K, epi):
    if M >= 40000 and epi not in (_lib.EPI_STORE_BF16, _lib.EPI_BIAS_RESID_F32):
        pytest.skip("large case covers the two memory-heaviest epilogues only")
    rng = np.random.default_rng(M + N + K + epi)
    A = dev(rng.standard_normal((M, K)), torch.bfloat16)
    W = rng.standard_normal((K, N)).astype(np.float32) / np.sqrt(K)       # flax kernel [in, out]
    Wt = dev(W.T, torch.bfloat16)                                         # packed [N, K]
    bias = dev(rng.standard_normal(N) * 0.5)
    resid = dev(rng.standard_normal((M, N)))
    acc = (A.double() @ Wt.double().t())                                  # fp64 on the same bf16 operands
    if epi == _lib.EPI_STORE_BF16:
        out = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
        want = acc
    elif epi == _lib.EPI_BIAS_GELU_BF16:
        out = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
        want = torch.nn.functional.gelu(acc + bias.double(), approximate="tanh")
    elif epi == _lib.EPI_BIAS_RESID_F32:
        out = resid.clone()
        want = acc + bias.double() + resid.double()
    else:
        out = torch.full((M, N), float("nan"), device="cuda")
        want = acc + bias.double()
    _lib.check(lib.vitb200_gemm_bf16(stream(), A.data_ptr(), Wt.data_ptr(), bias.data_ptr(),
                                     out.data_ptr(), M, N, K, epi, None, 0))
    torch.cuda.synchronize()
    err = (out.double() - want).abs().max().item()
    # fp32 accumulation: ~1e-5; bf16 outputs add one rounding (2^-9 relative, |values| < ~8);
    # tanh.approx in the GELU epilogue adds < 1e-3 absolute
    tol = 2e-4 if out.dtype == torch.float32 else 4e-2
    assert err < tol, f"max abs err {err}"
    if out.dtype == torch.bfloat16:                                       # and tight in the mean
        assert (out.double() - want).abs().mean().item() < 3e-3


def test_gemm_bf16_patch_epilogue(lib):
    # vit.py:147-153: Dense_0 output placed at row b*T+1+t with pos_embedding[1+t] added
    B, Np, K, D = 3, 16, 192, 64
    rng = np.random.default_rng(5)
    A = dev(rng.standard_normal((B * Np, K)), torch.bfloat16)
    Wt = dev(rng.standard_normal((D, K)) / np.sqrt(K), torch.bfloat16)
    bias = dev(rng.standard_normal(D))
    pos = dev(rng.standard_normal((Np + 1, D)))
    x = torch.full((B * (Np + 1), D), 7.0, device="cuda")
    _lib.check(lib.vitb200_gemm_bf16(stream(), A.data_ptr(), Wt.data_ptr(), bias.data_ptr(),
                                     x.data_ptr(), B * Np, D, K, _lib.EPI_PATCH_F32, pos.data_ptr(), Np))
    torch.cuda.synchronize()
    acc = (A.double() @ Wt.double().t()).view(B, Np, D) + bias.double() + pos.double()[1:]
    got = x.view(B, Np + 1, D)
    assert (got[:, 1:].double() - acc).abs().max().item() < 2e-4
    assert torch.all(got[:, 0] == 7.0)                                    # cls rows untouched


def test_gemm_bf16_rejects_unaligned(lib):
    a = torch.zeros((8, 12), dtype=torch.bfloat16, device="cuda")
    rc = lib.vitb200_gemm_bf16(stream(), a.data_ptr(), a.data_ptr(), None, a.data_ptr(), 8, 8, 12, 0, None, 0)
    assert rc == -1 and b"multiples of 8" in lib.vitb200_last_error()


@pytest.mark.parametrize("M,N,K", [(65, 1024, 3072), (130, 200, 77), (1, 1000, 1024)])
@pytest.mark.parametrize("epi", [_lib.EPI_STORE_BF16, _lib.EPI_BIAS_GELU_BF16, _lib.EPI_BIAS_RESID_F32, _lib.EPI_BIAS_F32])
def test_gemm_f32(lib, M, N, K, epi):
    rng = np.random.default_rng(M * 3 + N + K + epi)
    A = rng.standard_normal((M, K)); W = rng.standard_normal((K, N)) / np.sqrt(K)
    bias = rng.standard_normal(N); resid = rng.standard_normal((M, N))
    acc = A.astype(np.float32).astype(np.float64) @ W.astype(np.float32).astype(np.float64)
    b32 = bias.astype(np.float32).astype(np.float64)
    want = {0: acc, 1: vit_numpy.gelu_tanh(acc + b32), 2: acc + b32 + resid.astype(np.float32), 3: acc + b32}[epi]
    out = dev(resid) if epi == _lib.EPI_BIAS_RESID_F32 else torch.empty((M, N), device="cuda")
    _lib.check(lib.vitb200_gemm_f32(stream(), dev(A).data_ptr(), dev(W).data_ptr(), dev(bias).data_ptr(),
                                    out.data_ptr(), M, N, K, epi, None, 0))
    torch.cuda.synchronize()
    assert np.abs(out.cpu().numpy() - want).max() < 2e-5


@pytest.mark.parametrize("rows,dim", [(1000, 768), (197 * 4, 1024), (33, 1280), (64, 64), (9, 100), (5, 2052)])
@pytest.mark.parametrize("bf16", [0, 1])
def test_layernorm(lib, rows, dim, bf16):
    rng = np.random.default_rng(rows + dim)
    x = (rng.standard_normal((rows, dim)) * 2 + 3).astype(np.float32)
    p = {"scale": rng.standard_normal(dim).astype(np.float32), "bias": rng.standard_normal(dim).astype(np.float32)}
    want = vit_numpy.layer_norm(x.astype(np.float64), p)                  # eps 1e-6, vit.py:31
    y = torch.empty((rows, dim), dtype=torch.bfloat16 if bf16 else torch.float32, device="cuda")
    _lib.check(lib.vitb200_layernorm(stream(), dev(x).data_ptr(), dev(p["scale"]).data_ptr(),
                                     dev(p["bias"]).data_ptr(), y.data_ptr(), rows, dim, bf16))
    torch.cuda.synchronize()
    err = np.abs(y.float().cpu().numpy() - want).max()
    assert err < (4e-2 if bf16 else 2e-5)                                 # bf16: one rounding of |y| <~ 8


@pytest.mark.parametrize("batch,T,heads", [(2, 65, 16), (3, 197, 12), (2, 257, 16), (1, 1025, 4), (2, 16, 1), (1, 1, 2), (1, 130, 3)])
def test_attention_bf16(lib, batch, T, heads):
    rng = np.random.default_rng(T + heads)
    inner = heads * 64
    qkv = dev(rng.standard_normal((batch * T, 3 * inner)) * 1.5, torch.bfloat16)
    out = torch.empty((batch * T, inner), dtype=torch.bfloat16, device="cuda")
    _lib.check(lib.vitb200_attention_bf16(stream(), qkv.data_ptr(), out.data_ptr(), batch, T, heads))
    torch.cuda.synchronize()
    q, k, v = np.split(qkv.float().cpu().numpy().astype(np.float64).reshape(batch, T, 3 * inner), 3, axis=-1)
    th = lambda t: t.reshape(batch, T, heads, 64).transpose(0, 2, 1, 3)   # vit.py:71
    s = np.einsum("bhid,bhjd->bhij", th(q), th(k)) * 64 ** -0.5           # vit.py:73
    o = np.einsum("bhij,bhjd->bhid", vit_numpy.softmax_last(s), th(v))    # vit.py:75-78
    want = o.transpose(0, 2, 1, 3).reshape(batch * T, inner)              # vit.py:79
    err = np.abs(out.float().cpu().numpy() - want)
    # P is rounded to bf16 before PV (2^-9 relative) and the output to bf16: |o| <~ 4
    assert err.max() < 3e-2 and err.mean() < 3e-3, (err.max(), err.mean())


@pytest.mark.parametrize("batch,T,heads", [(1, 65, 16), (2, 197, 3), (1, 300, 2)])
def test_attention_f32(lib, batch, T, heads):
    rng = np.random.default_rng(T)
    inner = heads * 64
    qkv = rng.standard_normal((batch * T, 3 * inner)).astype(np.float32)
    out = torch.empty((batch * T, inner), device="cuda")
    _lib.check(lib.vitb200_attention_f32(stream(), dev(qkv).data_ptr(), out.data_ptr(), batch, T, heads))
    torch.cuda.synchronize()
    q, k, v = np.split(qkv.astype(np.float64).reshape(batch, T, 3 * inner), 3, axis=-1)
    th = lambda t: t.reshape(batch, T, heads, 64).transpose(0, 2, 1, 3)
    o = np.einsum("bhij,bhjd->bhid", vit_numpy.softmax_last(np.einsum("bhid,bhjd->bhij", th(q), th(k)) / 8), th(v))
    want = o.transpose(0, 2, 1, 3).reshape(batch * T, inner)
    assert np.abs(out.cpu().numpy() - want).max() < 1e-5


@pytest.mark.parametrize("H,W,ph,pw,Cc", [(224, 224, 16, 16, 3), (224, 224, 14, 14, 3), (16, 32, 8, 16, 3), (32, 32, 8, 8, 1)])
@pytest.mark.parametrize("bf16", [0, 1])
def test_patchify(lib, H, W, ph, pw, Cc, bf16):
    B = 2
    x = np.random.default_rng(H + pw).standard_normal((B, H, W, Cc)).astype(np.float32)
    K0 = ph * pw * Cc
    Kpad = (K0 + 63) // 64 * 64 if bf16 else K0 + (K0 & 1)
    Np = (H // ph) * (W // pw)
    out = torch.full((B * Np, Kpad), 9.0, dtype=torch.bfloat16 if bf16 else torch.float32, device="cuda")
    _lib.check(lib.vitb200_patchify(stream(), dev(x).data_ptr(), out.data_ptr(), B, H, W, Cc, ph, pw, Kpad, bf16))
    torch.cuda.synchronize()
    want = vit_numpy.patchify(x, ph, pw).reshape(B * Np, K0)              # vit.py:146
    got = out.float().cpu().numpy()
    ref = torch.as_tensor(want).to(torch.bfloat16).float().numpy() if bf16 else want
    np.testing.assert_array_equal(got[:, :K0], ref)                       # bit-exact (pure data movement + RN cast)
    assert not got[:, K0:].any()                                          # zero padding


def test_cls_rows_and_pool_layernorm(lib):
    B, T, D = 3, 17, 96
    rng = np.random.default_rng(1)
    cls, pos = rng.standard_normal(D).astype(np.float32), rng.standard_normal((T, D)).astype(np.float32)
    x = rng.standard_normal((B, T, D)).astype(np.float32)
    xd = dev(x)
    _lib.check(lib.vitb200_cls_rows(stream(), dev(cls).data_ptr(), dev(pos).data_ptr(), xd.data_ptr(), B, T, D))
    torch.cuda.synchronize()
    got = xd.cpu().numpy()
    np.testing.assert_array_equal(got[:, 0], np.broadcast_to(cls + pos[0], (B, D)))   # vit.py:151-153
    np.testing.assert_array_equal(got[:, 1:], x[:, 1:])
    p = {"scale": rng.standard_normal(D).astype(np.float32), "bias": rng.standard_normal(D).astype(np.float32)}
    for pool, name in ((0, "cls"), (1, "mean")):
        y = torch.empty((B, D), device="cuda")
        _lib.check(lib.vitb200_pool_layernorm(stream(), xd.data_ptr(), dev(p["scale"]).data_ptr(),
                                              dev(p["bias"]).data_ptr(), y.data_ptr(), B, T, D, pool, 0))
        torch.cuda.synchronize()
        g = got.astype(np.float64)
        pooled = g.mean(axis=1) if name == "mean" else g[:, 0]            # vit.py:159
        assert np.abs(y.cpu().numpy() - vit_numpy.layer_norm(pooled, p)).max() < 2e-5


def test_pack_weight(lib):
    K, N, Kpad = 588, 1280, 640                                           # ViT-H/14 patch kernel
    W = np.random.default_rng(2).standard_normal((K, N)).astype(np.float32)
    Wt = torch.full((N, Kpad), 3.0, dtype=torch.bfloat16, device="cuda")
    _lib.check(lib.vitb200_pack_weight_bf16(stream(), dev(W).data_ptr(), Wt.data_ptr(), K, N, Kpad))
    torch.cuda.synchronize()
    got = Wt.float().cpu().numpy()
    np.testing.assert_array_equal(got[:, :K], torch.as_tensor(W.T.copy()).to(torch.bfloat16).float().numpy())
    assert not got[:, K:].any()
