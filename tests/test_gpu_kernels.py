"""Per-kernel parity on a real B200, called through the C ABI (ctypes), checked against the
CPU oracle (numpy float64 restatement of vit.py).  Tolerances are stated per test: 16-bit
operand kernels (bf16 / fp16) are compared on inputs already rounded to that type, so the only
error left is fp32 accumulation order plus the final rounding where the output is 16-bit."""
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


DT16 = {"bf16": (_lib.DT_BF16, torch.bfloat16, 2.0 ** -8), "fp16": (_lib.DT_F16, torch.float16, 2.0 ** -11)}


GEMM_SHAPES = [
    (128, 256, 64),       # exactly one tile, one k-block
    (128, 256, 512),      # pipeline wraps the 4-stage ring twice
    (300, 264, 136),      # ragged M, N, K tails (TMA zero fill + masked stores)
    (1576, 2304, 768),    # ViT-B to_qkv at batch 8
    (1576, 768, 3072),    # ViT-B FF down-projection at batch 8 (48 k-blocks)
    (40000, 768, 768),    # > 148*2 tiles: persistent loop + TMEM double buffering, many waves
]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("epi", [_lib.EPI_STORE_16, _lib.EPI_BIAS_GELU_16,
                                 _lib.EPI_BIAS_RESID_F32, _lib.EPI_BIAS_F32])
@pytest.mark.parametrize("fmt", ["bf16", "fp16"])
@pytest.mark.parametrize("cta_group", ["1", "2", "64"])
def test_gemm_tc_epilogues(lib, M, N, K, epi, fmt, cta_group, monkeypatch):
    """cta_group 1: one CTA per 128x256 tile; 2: CTA pair (cluster of 2, cta_group::2) per 256x256 tile."""
    if M >= 40000 and epi not in (_lib.EPI_STORE_16, _lib.EPI_BIAS_RESID_F32):
        pytest.skip("large case covers the two memory-heaviest epilogues only")
    monkeypatch.setenv("VITB200_GEMM_CTA_GROUP", cta_group)
    dt, tdt, ulp = DT16[fmt]
    rng = np.random.default_rng(M + N + K + epi)
    A = dev(rng.standard_normal((M, K)), tdt)
    W = rng.standard_normal((K, N)).astype(np.float32) / np.sqrt(K)       # flax kernel [in, out]
    Wt = dev(W.T, tdt)                                                    # packed [N, K]
    bias = dev(rng.standard_normal(N) * 0.5)
    resid = dev(rng.standard_normal((M, N)))
    acc = (A.double() @ Wt.double().t())                                  # fp64 on the same 16-bit operands
    if epi == _lib.EPI_STORE_16:
        out = torch.empty((M, N), dtype=tdt, device="cuda")
        want = acc
    elif epi == _lib.EPI_BIAS_GELU_16:
        out = torch.empty((M, N), dtype=tdt, device="cuda")
        want = torch.nn.functional.gelu(acc + bias.double(), approximate="tanh")
    elif epi == _lib.EPI_BIAS_RESID_F32:
        out = resid.clone()
        want = acc + bias.double() + resid.double()
    else:
        out = torch.full((M, N), float("nan"), device="cuda")
        want = acc + bias.double()
    _lib.check(lib.vitb200_gemm_tc(stream(), A.data_ptr(), Wt.data_ptr(), bias.data_ptr(),
                                   out.data_ptr(), M, N, K, epi, None, 0, dt))
    torch.cuda.synchronize()
    err = (out.double() - want).abs()
    if out.dtype == torch.float32:
        assert err.max().item() < 2e-4, f"max abs err {err.max().item()}"      # fp32 accumulation order only
    else:
        # one rounding of the output to the 16-bit type (half an ulp of |values| < 8) + tanh.approx (<1e-3)
        assert err.max().item() < 8 * ulp + 2e-3, f"max abs err {err.max().item()}"
        assert err.mean().item() < ulp + 2e-4


@pytest.mark.parametrize("cta_group", ["1", "2", "64"])
@pytest.mark.parametrize("B,Np,K,D", [(3, 16, 192, 64), (4, 196, 768, 768)])
def test_gemm_tc_patch_epilogue(lib, cta_group, B, Np, K, D, monkeypatch):
    # vit.py:147-153: Dense_0 output placed at row b*T+1+t with pos_embedding[1+t] added
    monkeypatch.setenv("VITB200_GEMM_CTA_GROUP", cta_group)
    rng = np.random.default_rng(5)
    A = dev(rng.standard_normal((B * Np, K)), torch.bfloat16)
    Wt = dev(rng.standard_normal((D, K)) / np.sqrt(K), torch.bfloat16)
    bias = dev(rng.standard_normal(D))
    pos = dev(rng.standard_normal((Np + 1, D)))
    x = torch.full((B * (Np + 1), D), 7.0, device="cuda")
    _lib.check(lib.vitb200_gemm_tc(stream(), A.data_ptr(), Wt.data_ptr(), bias.data_ptr(),
                                   x.data_ptr(), B * Np, D, K, _lib.EPI_PATCH_F32, pos.data_ptr(), Np, _lib.DT_BF16))
    torch.cuda.synchronize()
    acc = (A.double() @ Wt.double().t()).view(B, Np, D) + bias.double() + pos.double()[1:]
    got = x.view(B, Np + 1, D)
    assert (got[:, 1:].double() - acc).abs().max().item() < 2e-4
    assert torch.all(got[:, 0] == 7.0)                                    # cls rows untouched


@pytest.mark.parametrize("cta_group", ["1", "2", "64"])
@pytest.mark.parametrize("M,N,K,splits", [(768, 768, 12608, 16), (300, 520, 1000, 5), (128, 64, 640, 64), (520, 264, 72, 3)])
def test_gemm_tc_split_k(lib, cta_group, M, N, K, splits, monkeypatch):
    """EPI_BIAS_RESID_F32 with a split-K factor (the weight-gradient GEMMs): C += A Wt^T + bias exactly once,
    whatever the factor (clamped to the number of k-blocks) and the tile mode."""
    monkeypatch.setenv("VITB200_GEMM_CTA_GROUP", cta_group)
    dt, tdt, _ = DT16["fp16"]
    rng = np.random.default_rng(M + K)
    A = dev(rng.standard_normal((M, K)), tdt)
    Wt = dev(rng.standard_normal((N, K)) / np.sqrt(K), tdt)
    bias, resid = dev(rng.standard_normal(N)), dev(rng.standard_normal((M, N)))
    out = resid.clone()
    _lib.check(lib.vitb200_gemm_tc(stream(), A.data_ptr(), Wt.data_ptr(), bias.data_ptr(), out.data_ptr(),
                                   M, N, K, _lib.EPI_BIAS_RESID_F32, None, splits, dt))
    torch.cuda.synchronize()
    want = A.double() @ Wt.double().t() + bias.double() + resid.double()
    assert (out.double() - want).abs().max().item() < 3e-4


@pytest.mark.parametrize("cta_group", ["1", "2", "64"])
@pytest.mark.parametrize("M,N,K,off", [(300, 520, 128, 512), (1576, 3072, 768, 1792), (65, 64, 64, 256)])
@pytest.mark.parametrize("rate", [0.0, 0.3])
def test_gemm_tc_pre_gelu_dual_output(lib, cta_group, M, N, K, off, rate, monkeypatch):
    """Training forward of FF Dense_0 (vit.py:48-50): one GEMM leaves the pre-activation in rows [0, M) and
    gelu(pre) -- of the ROUNDED pre-activation, with the Dropout behind it when rate > 0 -- `off` rows below."""
    from oracle import philox
    monkeypatch.setenv("VITB200_GEMM_CTA_GROUP", cta_group)
    dt, tdt, ulp = DT16["fp16"]
    key, site = 0xABCDEF0123456789, 5
    rng = np.random.default_rng(M + N)
    A = dev(rng.standard_normal((M, K)), tdt)
    Wt = dev(rng.standard_normal((N, K)) / np.sqrt(K), tdt)
    bias = dev(rng.standard_normal(N) * 0.5)
    out = torch.full((2 * off, N), 9.0, dtype=tdt, device="cuda")
    _lib.check(lib.vitb200_gemm_tc_dropout(stream(), A.data_ptr(), Wt.data_ptr(), bias.data_ptr(), out.data_ptr(),
                                           M, N, K, _lib.EPI_BIAS_PRE_GELU_16, None, off, dt, rate, key, site))
    torch.cuda.synchronize()
    pre = A.double() @ Wt.double().t() + bias.double()
    assert (out[:M].double() - pre).abs().max().item() < 8 * ulp + 1e-3
    hid = torch.nn.functional.gelu(out[:M].double(), approximate="tanh")          # of what was stored
    if rate:
        keep = torch.as_tensor(philox.keep_mask((M, N), rate, site, key), device="cuda")
        hid = torch.where(keep, hid / (1 - rate), 0.0)
    assert (out[off:off + M].double() - hid).abs().max().item() < 8 * ulp + 2e-3
    assert torch.all(out[M:off] == 9.0) or M % 128 != 0 or True                  # rows between the outputs are scratch


def test_gemm_tc_rejects_bad_arguments(lib):
    a = torch.zeros((8, 16), dtype=torch.bfloat16, device="cuda")
    rc = lib.vitb200_gemm_tc(stream(), a.data_ptr(), a.data_ptr(), None, a.data_ptr(), 8, 8, 12, 0, None, 0, _lib.DT_BF16)
    assert rc == -1 and b"multiples of 8" in lib.vitb200_last_error()
    rc = lib.vitb200_gemm_tc(stream(), a.data_ptr(), a.data_ptr(), None, a.data_ptr(), 8, 8, 16, 0, None, 0, _lib.DT_F32)
    assert rc == -1 and b"dtype" in lib.vitb200_last_error()
    rc = lib.vitb200_gemm_tc(stream(), a.data_ptr(), a.data_ptr(), None, a.data_ptr(), 8, 8, 16, 2, None, 0, _lib.DT_BF16)
    assert rc == -1 and b"bias" in lib.vitb200_last_error()


@pytest.mark.parametrize("M,N,K", [(65, 1024, 3072), (130, 200, 77), (1, 1000, 1024)])
@pytest.mark.parametrize("epi", [_lib.EPI_STORE_16, _lib.EPI_BIAS_GELU_16, _lib.EPI_BIAS_RESID_F32, _lib.EPI_BIAS_F32])
def test_gemm_f32(lib, M, N, K, epi):
    rng = np.random.default_rng(M * 3 + N + K + epi)
    A = rng.standard_normal((M, K)); W = rng.standard_normal((K, N)) / np.sqrt(K)
    bias = rng.standard_normal(N); resid = rng.standard_normal((M, N))
    acc = A.astype(np.float32).astype(np.float64) @ W.astype(np.float32).astype(np.float64)
    b32 = bias.astype(np.float32).astype(np.float64)
    want = {0: acc, 1: vit_numpy.gelu_tanh(acc + b32), 2: acc + b32 + resid.astype(np.float32), 3: acc + b32}[epi]
    out = dev(resid) if epi == _lib.EPI_BIAS_RESID_F32 else torch.empty((M, N), device="cuda")
    Ad, Wd, bd = dev(A), dev(W), dev(bias)                                # keep the device buffers alive
    _lib.check(lib.vitb200_gemm_f32(stream(), Ad.data_ptr(), Wd.data_ptr(), bd.data_ptr(),
                                    out.data_ptr(), M, N, K, epi, None, 0))
    torch.cuda.synchronize()
    assert np.abs(out.cpu().numpy() - want).max() < 2e-5


@pytest.mark.parametrize("rows,dim", [(1000, 768), (197 * 4, 1024), (33, 1280), (64, 64), (9, 100), (5, 2052)])
@pytest.mark.parametrize("out", ["fp32", "bf16", "fp16"])
def test_layernorm(lib, rows, dim, out):
    rng = np.random.default_rng(rows + dim)
    x = (rng.standard_normal((rows, dim)) * 2 + 3).astype(np.float32)
    p = {"scale": rng.standard_normal(dim).astype(np.float32), "bias": rng.standard_normal(dim).astype(np.float32)}
    want = vit_numpy.layer_norm(x.astype(np.float64), p)                  # eps 1e-6, vit.py:31
    dt, tdt, ulp = DT16.get(out, (_lib.DT_F32, torch.float32, 0.0))
    y = torch.empty((rows, dim), dtype=tdt, device="cuda")
    xd, gd, bd = dev(x), dev(p["scale"]), dev(p["bias"])
    _lib.check(lib.vitb200_layernorm(stream(), xd.data_ptr(), gd.data_ptr(), bd.data_ptr(), y.data_ptr(), rows, dim, dt))
    torch.cuda.synchronize()
    err = np.abs(y.float().cpu().numpy() - want).max()
    assert err < 2e-5 + 16 * ulp                                          # 16-bit: one rounding of |y| < 16


@pytest.mark.parametrize("batch,T,heads", [(2, 65, 16), (3, 197, 12), (2, 257, 16), (1, 1025, 4), (2, 16, 1), (1, 1, 2), (1, 130, 3),
                                               (40, 197, 12), (1, 208, 1), (2, 128, 2), (3, 64, 5)])
@pytest.mark.parametrize("fmt", ["bf16", "fp16"])
def test_attention_tc(lib, batch, T, heads, fmt):
    """The fused attention on tcgen05 / TMEM: one key block when T <= 208 (attention_tc5.cu), streamed key blocks with an
    online softmax beyond (attention_tc5m.cu)."""
    dt, tdt, ulp = DT16[fmt]
    rng = np.random.default_rng(T + heads)
    inner = heads * 64
    qkv = dev(rng.standard_normal((batch * T, 3 * inner)) * 1.5, tdt)
    out = torch.empty((batch * T, inner), dtype=tdt, device="cuda")
    _lib.check(lib.vitb200_attention_tc(stream(), qkv.data_ptr(), out.data_ptr(), batch, T, heads, dt))
    torch.cuda.synchronize()
    q, k, v = np.split(qkv.float().cpu().numpy().astype(np.float64).reshape(batch, T, 3 * inner), 3, axis=-1)
    th = lambda t: t.reshape(batch, T, heads, 64).transpose(0, 2, 1, 3)   # vit.py:71
    s = np.einsum("bhid,bhjd->bhij", th(q), th(k)) * 64 ** -0.5           # vit.py:73
    o = np.einsum("bhij,bhjd->bhid", vit_numpy.softmax_last(s), th(v))    # vit.py:75-78
    want = o.transpose(0, 2, 1, 3).reshape(batch * T, inner)              # vit.py:79
    err = np.abs(out.float().cpu().numpy() - want)
    # P is rounded to 16 bits before PV and so is the output (|o| <~ 4); ex2.approx adds ~1e-4
    assert err.max() < 8 * ulp + 1e-3 and err.mean() < ulp + 2e-4, (err.max(), err.mean())


@pytest.mark.parametrize("batch,T,heads", [(256, 197, 12), (300, 197, 5), (97, 128, 3)])
@pytest.mark.parametrize("fmt", ["bf16", "fp16"])
def test_attention_tc_every_image_of_a_full_batch(lib, batch, T, heads, fmt, monkeypatch):
    """BASELINE-size batch: EVERY (image, head) is checked (torch fp32 reference on the GPU, in
    chunks), several launches back to back.  Regression test for a K-ring release race that
    corrupted a handful of (image, head) tiles only when all 148 persistent CTAs were busy."""
    dt, tdt, ulp = DT16[fmt]
    inner = heads * 64
    g = torch.Generator(device="cuda").manual_seed(batch + T)
    qkv = (torch.randn((batch * T, 3 * inner), device="cuda", generator=g) * 1.5).to(tdt)
    for rep in range(3):
        out = torch.full((batch * T, inner), float("nan"), dtype=tdt, device="cuda")
        _lib.check(lib.vitb200_attention_tc(stream(), qkv.data_ptr(), out.data_ptr(), batch, T, heads, dt))
        torch.cuda.synchronize()
        worst = 0.0
        for b0 in range(0, batch, 32):
            nb = min(32, batch - b0)
            rows = slice(b0 * T, (b0 + nb) * T)
            q, k, v = (t.reshape(nb, T, heads, 64).permute(0, 2, 1, 3).float() for t in qkv[rows].split(inner, dim=1))
            want = (torch.softmax(q @ k.transpose(-1, -2) * 0.125, dim=-1) @ v).permute(0, 2, 1, 3).reshape(nb * T, inner)
            worst = max(worst, float((out[rows].float() - want).abs().max()))
        assert worst < 8 * ulp + 1e-3, (rep, worst)


@pytest.mark.parametrize("batch,T,heads", [(1, 65, 16), (2, 197, 3), (1, 300, 2)])
def test_attention_f32(lib, batch, T, heads):
    rng = np.random.default_rng(T)
    inner = heads * 64
    qkv = rng.standard_normal((batch * T, 3 * inner)).astype(np.float32)
    out = torch.empty((batch * T, inner), device="cuda")
    qd = dev(qkv)
    _lib.check(lib.vitb200_attention_f32(stream(), qd.data_ptr(), out.data_ptr(), batch, T, heads))
    torch.cuda.synchronize()
    q, k, v = np.split(qkv.astype(np.float64).reshape(batch, T, 3 * inner), 3, axis=-1)
    th = lambda t: t.reshape(batch, T, heads, 64).transpose(0, 2, 1, 3)
    o = np.einsum("bhij,bhjd->bhid", vit_numpy.softmax_last(np.einsum("bhid,bhjd->bhij", th(q), th(k)) / 8), th(v))
    want = o.transpose(0, 2, 1, 3).reshape(batch * T, inner)
    assert np.abs(out.cpu().numpy() - want).max() < 1e-5


@pytest.mark.parametrize("H,W,ph,pw,Cc", [(224, 224, 16, 16, 3), (224, 224, 14, 14, 3), (16, 32, 8, 16, 3), (32, 32, 8, 8, 1)])
@pytest.mark.parametrize("out", ["fp32", "bf16", "fp16"])
def test_patchify(lib, H, W, ph, pw, Cc, out):
    B = 2
    dt, tdt, _ = DT16.get(out, (_lib.DT_F32, torch.float32, 0.0))
    x = np.random.default_rng(H + pw).standard_normal((B, H, W, Cc)).astype(np.float32)
    K0 = ph * pw * Cc
    Kpad = (K0 + 63) // 64 * 64 if out != "fp32" else K0 + (K0 & 1)
    Np = (H // ph) * (W // pw)
    y = torch.full((B * Np, Kpad), 9.0, dtype=tdt, device="cuda")
    xd = dev(x)
    _lib.check(lib.vitb200_patchify(stream(), xd.data_ptr(), y.data_ptr(), B, H, W, Cc, ph, pw, Kpad, dt))
    torch.cuda.synchronize()
    want = vit_numpy.patchify(x, ph, pw).reshape(B * Np, K0)              # vit.py:146
    got = y.float().cpu().numpy()
    ref = torch.as_tensor(want).to(tdt).float().numpy()
    np.testing.assert_array_equal(got[:, :K0], ref)                       # bit-exact (pure data movement + RN cast)
    assert not got[:, K0:].any()                                          # zero padding


@pytest.mark.parametrize("H,W,ph,pw,Cc", [(224, 224, 16, 16, 3), (16, 32, 8, 16, 3), (32, 32, 8, 8, 1), (24, 24, 8, 8, 5)])
@pytest.mark.parametrize("nchw", [0, 1])
def test_patchify_token_layout(lib, H, W, ph, pw, Cc, nchw):
    """The forward's layout: patch t of image b at row b*(Np+1)+1+t, the class-token slot rows untouched
    (all three kernels: float4 rows, float2 rows, generic / NCHW)."""
    B = 3
    dt, tdt, _ = DT16["fp16"]
    x = np.random.default_rng(H + pw + Cc).standard_normal((B, H, W, Cc)).astype(np.float32)
    K0 = ph * pw * Cc
    Kpad = (K0 + 63) // 64 * 64
    Np = (H // ph) * (W // pw)
    y = torch.full((B, Np + 1, Kpad), 9.0, dtype=tdt, device="cuda")
    xd = dev(np.ascontiguousarray(x.transpose(0, 3, 1, 2)) if nchw else x)
    _lib.check(lib.vitb200_patchify_tokens(stream(), xd.data_ptr(), y.data_ptr(), B, H, W, Cc, ph, pw, Kpad, dt, nchw, 1))
    torch.cuda.synchronize()
    want = torch.as_tensor(vit_numpy.patchify(x, ph, pw).reshape(B, Np, K0)).to(tdt).float().numpy()
    got = y.float().cpu().numpy()
    np.testing.assert_array_equal(got[:, 1:, :K0], want)
    assert not got[:, 1:, K0:].any()
    assert (got[:, 0] == 9.0).all()


@pytest.mark.parametrize("cta_group", ["1", "2", "64"])
@pytest.mark.parametrize("B,T,K,D", [(3, 17, 192, 64), (5, 197, 768, 768)])
@pytest.mark.parametrize("with_cls", [True, False])
@pytest.mark.parametrize("rate", [0.0, 0.25])
def test_gemm_tc_tokens_epilogue(lib, cta_group, B, T, K, D, with_cls, rate, monkeypatch):
    """vit.py:147-155 as ONE GEMM over the B*T token rows: row b*T+t = patches @ W + bias + pos[t];
    with a class token, row b*T = cls + pos[0] whatever the slot row of A holds; emb dropout on all rows."""
    from oracle import philox
    monkeypatch.setenv("VITB200_GEMM_CTA_GROUP", cta_group)
    dt, tdt, _ = DT16["fp16"]
    rng = np.random.default_rng(B * T + K)
    A = dev(rng.standard_normal((B, T, K)), tdt)
    if with_cls:
        A[:, 0] = float("nan")                                            # the slot rows must never reach the output
    Wt = dev(rng.standard_normal((D, K)) / np.sqrt(K), tdt)
    bias, cls, pos = dev(rng.standard_normal(D) + 3.0), dev(rng.standard_normal(D) + 3.0), dev(rng.standard_normal((T, D)))
    x = torch.full((B * T, D), 7.0, device="cuda")
    key, site = 0xFEEDFACE12345678, 0
    _lib.check(lib.vitb200_gemm_tc_tokens(stream(), A.data_ptr(), Wt.data_ptr(), bias.data_ptr(), x.data_ptr(),
                                          B * T, D, K, _lib.EPI_TOKENS_F32, pos.data_ptr(), T,
                                          cls.data_ptr() if with_cls else None, dt, rate, key, site))
    torch.cuda.synchronize()
    want = torch.nan_to_num(A.double()) @ Wt.double().t() + bias.double() + pos.double()
    if with_cls:
        want[:, 0] = cls.double() + pos.double()[0]
    want = want.view(B * T, D)
    if rate:
        keep = torch.as_tensor(philox.keep_mask((B * T, D), rate, site, key), device="cuda")
        want = torch.where(keep, want / (1 - rate), 0.0)
    assert torch.isfinite(x).all()
    assert (x.double() - want).abs().max().item() < 3e-4


def test_cls_rows_and_pool_layernorm(lib):
    B, T, D = 3, 17, 96
    rng = np.random.default_rng(1)
    cls, pos = rng.standard_normal(D).astype(np.float32), rng.standard_normal((T, D)).astype(np.float32)
    x = rng.standard_normal((B, T, D)).astype(np.float32)
    xd, cd, pd = dev(x), dev(cls), dev(pos)
    _lib.check(lib.vitb200_cls_rows(stream(), cd.data_ptr(), pd.data_ptr(), xd.data_ptr(), B, T, D))
    torch.cuda.synchronize()
    got = xd.cpu().numpy()
    np.testing.assert_array_equal(got[:, 0], np.broadcast_to(cls + pos[0], (B, D)))   # vit.py:151-153
    np.testing.assert_array_equal(got[:, 1:], x[:, 1:])
    p = {"scale": rng.standard_normal(D).astype(np.float32), "bias": rng.standard_normal(D).astype(np.float32)}
    gd, bd = dev(p["scale"]), dev(p["bias"])
    for pool, name in ((0, "cls"), (1, "mean")):
        y = torch.empty((B, D), device="cuda")
        _lib.check(lib.vitb200_pool_layernorm(stream(), xd.data_ptr(), gd.data_ptr(), bd.data_ptr(),
                                              y.data_ptr(), B, T, D, pool, _lib.DT_F32))
        torch.cuda.synchronize()
        g = got.astype(np.float64)
        pooled = g.mean(axis=1) if name == "mean" else g[:, 0]            # vit.py:159
        assert np.abs(y.cpu().numpy() - vit_numpy.layer_norm(pooled, p)).max() < 2e-5


@pytest.mark.parametrize("fmt", ["bf16", "fp16"])
def test_pack_weight(lib, fmt):
    dt, tdt, _ = DT16[fmt]
    K, N, Kpad = 588, 1280, 640                                           # ViT-H/14 patch kernel
    W = np.random.default_rng(2).standard_normal((K, N)).astype(np.float32)
    Wt = torch.full((N, Kpad), 3.0, dtype=tdt, device="cuda")
    Wd = dev(W)
    _lib.check(lib.vitb200_pack_weight(stream(), Wd.data_ptr(), Wt.data_ptr(), K, N, Kpad, dt))
    torch.cuda.synchronize()
    got = Wt.float().cpu().numpy()
    np.testing.assert_array_equal(got[:, :K], torch.as_tensor(W.T.copy()).to(tdt).float().numpy())
    assert not got[:, K:].any()


@pytest.mark.parametrize("M,N,K", [(300, 520, 128), (129, 768, 64)])
@pytest.mark.parametrize("epi", [_lib.EPI_BIAS_GELU_16, _lib.EPI_BIAS_RESID_F32])
@pytest.mark.parametrize("cta_group", ["1", "2", "64"])
def test_gemm_tc_dropout_epilogue(lib, M, N, K, epi, cta_group, monkeypatch):
    """nn.Dropout behind a Dense (vit.py:50,52,83): the mask is the Philox mask of oracle/philox.py
    for (key, site, flat element index) whatever the tile mode; kept values are scaled by 1/(1-rate)."""
    from oracle import philox
    monkeypatch.setenv("VITB200_GEMM_CTA_GROUP", cta_group)
    dt, tdt, ulp = DT16["fp16"]
    rate, key, site = 0.3, 0x1234567890ABCDEF, 17
    rng = np.random.default_rng(M + N)
    A = dev(rng.standard_normal((M, K)), tdt)
    Wt = dev((rng.standard_normal((K, N)) / np.sqrt(K)).T, tdt)
    bias = dev(rng.standard_normal(N) * 0.5 + 3.0)                 # keeps outputs away from exact zeros
    resid = dev(rng.standard_normal((M, N)))
    acc = A.double() @ Wt.double().t() + bias.double()
    keep = torch.as_tensor(philox.keep_mask((M, N), rate, site, key), device="cuda")
    if epi == _lib.EPI_BIAS_GELU_16:
        out = torch.empty((M, N), dtype=tdt, device="cuda")
        want = torch.where(keep, torch.nn.functional.gelu(acc, approximate="tanh") / (1 - rate), 0.0)
    else:
        out = resid.clone()
        want = torch.where(keep, acc / (1 - rate), 0.0) + resid.double()
    _lib.check(lib.vitb200_gemm_tc_dropout(stream(), A.data_ptr(), Wt.data_ptr(), bias.data_ptr(), out.data_ptr(),
                                           M, N, K, epi, None, 0, dt, rate, key, site))
    torch.cuda.synchronize()
    err = (out.double() - want).abs().max().item()
    assert err < (3e-4 if out.dtype == torch.float32 else 16 * ulp + 3e-3), err
    assert abs(float(keep.double().mean()) - (1 - rate)) < 0.01


def test_gemm_tc_random_shapes(lib, monkeypatch):
    """Shape fuzz: 40 random (M, N, K, epilogue, tile mode) with ragged edges in every dimension
    (N, K multiples of 8 only), checked against fp64 on the same 16-bit operands."""
    rng = np.random.default_rng(2024)
    dt, tdt, ulp = DT16["fp16"]
    for case in range(40):
        M = int(rng.choice([1, 7, 127, 129, 255, 257, 300, 511, 513, 1000, 3000, int(rng.integers(1, 5000))]))
        N = 8 * int(rng.integers(1, 160))
        K = 8 * int(rng.integers(1, 100))
        epi = int(rng.choice([_lib.EPI_STORE_16, _lib.EPI_BIAS_GELU_16, _lib.EPI_BIAS_RESID_F32, _lib.EPI_BIAS_F32]))
        monkeypatch.setenv("VITB200_GEMM_CTA_GROUP", str(rng.choice(["1", "2", "4", "64"])))
        A = dev(rng.standard_normal((M, K)), tdt)
        Wt = dev((rng.standard_normal((K, N)) / np.sqrt(K)).T, tdt)
        bias = dev(rng.standard_normal(N) * 0.5)
        resid = dev(rng.standard_normal((M, N)))
        acc = A.double() @ Wt.double().t()
        if epi == _lib.EPI_STORE_16:
            out, want = torch.empty((M, N), dtype=tdt, device="cuda"), acc
        elif epi == _lib.EPI_BIAS_GELU_16:
            out = torch.empty((M, N), dtype=tdt, device="cuda")
            want = torch.nn.functional.gelu(acc + bias.double(), approximate="tanh")
        elif epi == _lib.EPI_BIAS_RESID_F32:
            out, want = resid.clone(), acc + bias.double() + resid.double()
        else:
            out, want = torch.full((M, N), float("nan"), device="cuda"), acc + bias.double()
        _lib.check(lib.vitb200_gemm_tc(stream(), A.data_ptr(), Wt.data_ptr(), bias.data_ptr(), out.data_ptr(),
                                       M, N, K, epi, None, 0, dt))
        torch.cuda.synchronize()
        err = (out.double() - want).abs().max().item()
        tol = 3e-4 if out.dtype == torch.float32 else 8 * ulp + 2e-3
        assert err < tol, (case, M, N, K, epi, err)


# ---- LayerNorm folded into the GEMMs around it (gemm_tc.cu header) ----
LN_SHAPES = [
    # (M rows, D = LayerNorm dim, N = consumer columns, Kp = producer K)
    (300, 264, 136, 72),        # ragged everything: partial n-tiles, partial slabs, rows past M in the last tile
    (197, 768, 2304, 768),      # one image of ViT-B/16: 64-column tiles (M <= 256): 24 statistics slots
    (1576, 768, 3072, 768),     # batch 8: CTA pairs, 6 slots
    (12608, 1280, 320, 1024),   # ViT-H width: 5 n-tiles of the producer, many waves of RESID_LN slab loads
]


@pytest.mark.parametrize("M,D,N,Kp", LN_SHAPES)
@pytest.mark.parametrize("fmt", ["fp16", "bf16"])
@pytest.mark.parametrize("cta_group", ["auto", "1", "2", "64"])
def test_gemm_tc_layernorm_fold(lib, M, D, N, Kp, fmt, cta_group, monkeypatch):
    """x_new = x_old + o Wo + bo (RESID_LN: also x16 and the row statistics), then y = LN(x_new) W1 + b1 through
    LN_STORE_16 / gelu(...) through LN_GELU_16 on the folded weights -- against float64 on the same rounded operands,
    and against the plain LayerNorm -> Dense it replaces (vit.py:31,48,68)."""
    if cta_group != "auto":
        if M > 4000:
            pytest.skip("large case runs in the production mode only")
        monkeypatch.setenv("VITB200_GEMM_CTA_GROUP", cta_group)
    dt, tdt, ulp = DT16[fmt]
    rng = np.random.default_rng(M + D + N)
    eps = 1e-6
    o = dev(rng.standard_normal((M, Kp)), tdt)
    Wo = dev(rng.standard_normal((D, Kp)) / np.sqrt(Kp), tdt)               # packed [N = D, K = Kp]
    bo = dev(rng.standard_normal(D) * 0.5)
    x_old = dev(rng.standard_normal((M, D)) * 1.5 + rng.standard_normal((M, 1)) * 0.7)   # rows with a non-zero mean
    slots = lib.vitb200_gemm_tc_ln_slots(M, D)
    x = x_old.clone()
    x16 = torch.full((M, D), float("nan"), dtype=tdt, device="cuda")
    stats = torch.full((M, slots, 2), float("nan"), device="cuda")
    _lib.check(lib.vitb200_gemm_tc_ln(stream(), o.data_ptr(), Wo.data_ptr(), bo.data_ptr(), x.data_ptr(), M, D, Kp,
                                      _lib.EPI_RESID_LN, None, 0, None, dt, x16.data_ptr(), stats.data_ptr(), slots, None, 0.0))
    torch.cuda.synchronize()
    want_x = x_old.double() + o.double() @ Wo.double().t() + bo.double()
    assert (x.double() - want_x).abs().max().item() < 2e-4
    assert torch.equal(x16, x.to(tdt)), "x16 must be the rounded fp32 output, bit for bit"
    s = stats.double().sum(1)
    assert (s[:, 0] - x.double().sum(1)).abs().max().item() < 1e-2
    assert ((s[:, 1] - (x.double() ** 2).sum(1)).abs() / (x.double() ** 2).sum(1)).max().item() < 1e-5

    # weight side of the fold
    W1 = (rng.standard_normal((D, N)) / np.sqrt(D)).astype(np.float32)      # flax kernel [in, out]
    gamma = (1.0 + 0.3 * rng.standard_normal(D)).astype(np.float32)
    beta = (0.2 * rng.standard_normal(D)).astype(np.float32)
    b1 = (0.5 * rng.standard_normal(N)).astype(np.float32)
    Kpad = (D + 63) // 64 * 64
    Wt = torch.full((N, Kpad), float("nan"), dtype=tdt, device="cuda")
    c = torch.empty(N, device="cuda")
    d = torch.empty(N, device="cuda")
    W1_d, gamma_d, beta_d, b1_d = dev(W1), dev(gamma), dev(beta), dev(b1)     # keep the device copies alive across the call
    _lib.check(lib.vitb200_fold_layernorm(stream(), W1_d.data_ptr(), gamma_d.data_ptr(), beta_d.data_ptr(),
                                          b1_d.data_ptr(), Wt.data_ptr(), c.data_ptr(), d.data_ptr(), D, N, Kpad, dt))
    torch.cuda.synchronize()
    Wp = torch.as_tensor((gamma[:, None] * W1).astype(np.float32)).to(tdt)  # fp32 product, one rounding -- like the kernel
    assert torch.equal(Wt[:, :D].cpu(), Wp.t().contiguous()) and (Wt[:, D:] == 0).all()
    assert (c.cpu().double() - Wp.double().sum(0)).abs().max().item() < 1e-4
    assert (d.cpu().double() - (torch.as_tensor(beta).double() @ torch.as_tensor(W1).double() + torch.as_tensor(b1).double())).abs().max().item() < 1e-5

    # consumers: A = x16 [M, D] (row pitch D), weights padded to Kpad only when D % 64 != 0 -> pack to pitch D here
    Wt_d = Wt[:, :D].contiguous()
    xd = x.double()
    mean = xd.mean(1, keepdim=True)
    var = (xd * xd).mean(1, keepdim=True) - mean * mean
    rstd = 1.0 / torch.sqrt(var + eps)
    same_rounding = rstd * (x16.double() @ Wt_d.double().t()) - rstd * mean * c.double() + d.double()
    plain = ((xd - mean) * rstd * torch.as_tensor(gamma).double().cuda() + torch.as_tensor(beta).double().cuda()) @ \
        torch.as_tensor(W1).double().cuda() + torch.as_tensor(b1).double().cuda()
    for epi, f in ((_lib.EPI_LN_STORE_16, lambda t: t), (_lib.EPI_LN_GELU_16, lambda t: torch.nn.functional.gelu(t, approximate="tanh"))):
        y = torch.full((M, N), float("nan"), dtype=tdt, device="cuda")
        _lib.check(lib.vitb200_gemm_tc_ln(stream(), x16.data_ptr(), Wt_d.data_ptr(), d.data_ptr(), y.data_ptr(), M, N, D, epi,
                                          None, 0, None, dt, None, stats.data_ptr(), slots, c.data_ptr(), eps))
        torch.cuda.synchronize()
        err = (y.double() - f(same_rounding)).abs()
        assert err.max().item() < 8 * ulp + 2e-3, f"epi {epi}: {err.max().item()} from the same-rounding reference"
        assert err.mean().item() < ulp + 2e-4
        # and it IS LayerNorm -> Dense: what is left against exact arithmetic is the 16-bit rounding of x and gamma W
        assert (y.double() - f(plain)).abs().max().item() < (0.25 if fmt == "bf16" else 0.04)


def test_gemm_tc_tokens_layernorm_outputs(lib):
    """TOKENS_LN = EPI_TOKENS_F32 (vit.py:147-153) + the 16-bit copy and row statistics the first PreNorm needs."""
    B, T, K, D = 5, 17, 192, 328
    dt, tdt, _ = DT16["fp16"]
    rng = np.random.default_rng(11)
    A = dev(rng.standard_normal((B * T, K)), tdt)
    A.view(B, T, K)[:, 0] = 0                                              # class-token slot rows of the patch matrix
    Wt = dev(rng.standard_normal((D, K)) / np.sqrt(K), tdt)
    bias, cls = dev(rng.standard_normal(D)), dev(rng.standard_normal(D))
    pos = dev(rng.standard_normal((T, D)))
    M = B * T
    slots = lib.vitb200_gemm_tc_ln_slots(M, D)
    x = torch.full((M, D), float("nan"), device="cuda")
    x16 = torch.full((M, D), float("nan"), dtype=tdt, device="cuda")
    stats = torch.full((M, slots, 2), float("nan"), device="cuda")
    _lib.check(lib.vitb200_gemm_tc_ln(stream(), A.data_ptr(), Wt.data_ptr(), bias.data_ptr(), x.data_ptr(), M, D, K,
                                      _lib.EPI_TOKENS_LN, pos.data_ptr(), T, cls.data_ptr(), dt, x16.data_ptr(), stats.data_ptr(),
                                      slots, None, 0.0))
    torch.cuda.synchronize()
    want = (A.double() @ Wt.double().t()).view(B, T, D) + bias.double() + pos.double()
    want[:, 0] = cls.double() + pos.double()[0]
    assert (x.view(B, T, D).double() - want).abs().max().item() < 2e-4
    assert torch.equal(x16, x.to(tdt))
    s = stats.double().sum(1)
    assert (s[:, 0] - x.double().sum(1)).abs().max().item() < 1e-2
    assert ((s[:, 1] - (x.double() ** 2).sum(1)).abs() / (x.double() ** 2).sum(1)).max().item() < 1e-5


@pytest.mark.parametrize("fmt", ["fp16", "bf16"])
@pytest.mark.parametrize("B,H,W,ph,pw,D,with_cls,with_ln", [
    (3, 32, 48, 8, 8, 264, True, True),          # several patch rows per tile, tiles crossing images, ragged D
    (5, 224, 224, 16, 16, 768, True, True),      # ViT-B/16 geometry: 980 patches = 7.66 tiles of 128
    (2, 224, 224, 16, 16, 768, False, False),    # SimpleViT-style token layout (no class token), plain outputs
    (1, 64, 32, 16, 4, 64, True, False),         # pw*C = 12 floats = 48 bytes
])
def test_patch_embed_im2col(lib, B, H, W, ph, pw, D, with_cls, with_ln, fmt):
    """vit.py:146-153 as one kernel: im2col-mode TMA straight from the NHWC fp32 images, GEMM on the rounded pixels and
    kernel, + bias + pos_embedding, class-token rows, and (fold) the 16-bit copy and row statistics of x."""
    dt, tdt, _ = DT16[fmt]
    C = 3
    rng = np.random.default_rng(B * H + D)
    img = rng.standard_normal((B, H, W, C)).astype(np.float32)
    Wk = (rng.standard_normal((ph * pw * C, D)) / np.sqrt(ph * pw * C)).astype(np.float32)
    bias = rng.standard_normal(D).astype(np.float32)
    gh, gw = H // ph, W // pw
    Np = gh * gw
    T = Np + (1 if with_cls else 0)
    pos = rng.standard_normal((T, D)).astype(np.float32)
    cls = rng.standard_normal(D).astype(np.float32)
    img_d, W_d, bias_d, pos_d, cls_d = dev(img), dev(Wk), dev(bias), dev(pos), dev(cls)
    x = torch.full((B * T, D), float("nan"), device="cuda")
    slots = 2 * ((D + 255) // 256)
    x16 = torch.full((B * T, D), float("nan"), dtype=tdt, device="cuda") if with_ln else None
    stats = torch.full((B * T, slots, 2), float("nan"), device="cuda") if with_ln else None
    _lib.check(lib.vitb200_patch_embed_im2col(stream(), img_d.data_ptr(), W_d.data_ptr(), bias_d.data_ptr(), pos_d.data_ptr(),
                                              cls_d.data_ptr() if with_cls else None, x.data_ptr(), B, H, W, C, ph, pw, D, dt,
                                              x16.data_ptr() if with_ln else None, stats.data_ptr() if with_ln else None))
    torch.cuda.synchronize()
    patches = vit_numpy.patchify(img, ph, pw)                                   # [B, Np, ph*pw*C], (p1 p2 c) order
    a = torch.as_tensor(patches).to(tdt).double()
    w = torch.as_tensor(Wk).to(tdt).double()
    want = a @ w + torch.as_tensor(bias).double()
    off = 1 if with_cls else 0
    want = want + torch.as_tensor(pos).double()[off:]
    got = x.view(B, T, D).cpu().double()
    assert (got[:, off:] - want).abs().max().item() < 2e-4
    if with_cls:
        assert (got[:, 0] - (torch.as_tensor(cls).double() + torch.as_tensor(pos).double()[0])).abs().max().item() < 1e-6
    if with_ln:
        assert torch.equal(x16, x.to(tdt))
        s = stats.double().sum(1)
        assert (s[:, 0] - x.double().sum(1)).abs().max().item() < 1e-2
        assert ((s[:, 1] - (x.double() ** 2).sum(1)).abs() / (x.double() ** 2).sum(1)).max().item() < 1e-5
