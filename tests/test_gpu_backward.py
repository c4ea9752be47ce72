"""Backward pass on a real B200 against autograd through the CPU oracle.

The reference never differentiates its forward, so the checker is ``oracle/vit_torch.vit_vjp``:
float64 torch autograd through the restatement of vit.py (itself pinned by tests/test_oracle.py,
including a finite-difference check of this function).  The CUDA path runs 16-bit GEMM operands
and keeps 16-bit activations / activation gradients, so gradients are compared leaf by leaf as
max|g - g_ref| / max|g_ref| with the tolerance stated in each test."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import vit_torch
from vit_flax_b200 import ViT, _lib, init_params, perturb_params
from vit_flax_b200._lib import VitB200Error
from vit_flax_b200.engine import Engine
from vit_flax_b200.params import flatten_params
from vit_flax_b200.runtime import clear_cache
from _util import C2, TINY, images_for

pytestmark = pytest.mark.gpu

DT16 = {"bf16": (_lib.DT_BF16, torch.bfloat16), "fp16": (_lib.DT_F16, torch.float16)}


@pytest.fixture(scope="module")
def lib(lib_built):
    assert lib_built.vitb200_device_count() >= 1, "no sm_100 device"
    return lib_built


@pytest.fixture(autouse=True)
def _fresh_cache():
    yield
    clear_cache()
    torch.cuda.empty_cache()


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def rel_err(got, want):
    want = np.asarray(want, dtype=np.float64)
    return float(np.abs(np.asarray(got, dtype=np.float64) - want).max() / max(np.abs(want).max(), 1e-30))


@pytest.mark.parametrize("batch,T,heads", [(2, 197, 3), (1, 208, 1), (3, 64, 2), (2, 17, 2), (1, 100, 1), (40, 197, 12),
                                           (2, 257, 2), (1, 1025, 1), (3, 300, 2), (2, 209, 1),
                                           (30, 257, 8), (5, 1025, 7), (150, 128, 3)])   # several units of work per CTA
@pytest.mark.parametrize("fmt", ["fp16", "bf16"])
def test_attention_bwd(lib, batch, T, heads, fmt):
    """Adjoint of vit.py:69-79 per (image, head): dq, dk, dv from (q, k, v, d_out), all 16-bit, on tcgen05
    (attention_bwd_tc5.cu): resident form up to T = 208, streamed form (dQ summed in an fp32 buffer) beyond.  This per-kernel
    entry point has no log-sum-exp from a forward pass: it re-runs the forward kernel for it (the model path keeps it)."""
    dt, tdt = DT16[fmt]
    inner = heads * 64
    g = torch.Generator().manual_seed(T * 7 + heads)
    qkv = (torch.randn((batch * T, 3 * inner), generator=g) * 1.2).to(tdt).cuda()
    d_out = torch.randn((batch * T, inner), generator=g).to(tdt).cuda()
    dqkv = torch.full((batch * T, 3 * inner), 9.0, dtype=tdt, device="cuda")
    out = torch.empty((batch * T, inner), dtype=tdt, device="cuda")             # the forward's output, as train_forward keeps it
    _lib.check(lib.vitb200_attention_tc(stream(), qkv.data_ptr(), out.data_ptr(), batch, T, heads, dt))
    _lib.check(lib.vitb200_attention_bwd(stream(), qkv.data_ptr(), out.data_ptr(), d_out.data_ptr(), dqkv.data_ptr(), batch, T, heads, dt))
    torch.cuda.synchronize()
    x = qkv.double().view(batch, T, 3, heads, 64).permute(2, 0, 3, 1, 4).contiguous().requires_grad_(True)   # [3, B, h, T, 64]
    o = torch.softmax(x[0] @ x[1].transpose(-1, -2) / 8.0, dim=-1) @ x[2]                                      # vit.py:73-78
    (o * d_out.double().view(batch, T, heads, 64).permute(0, 2, 1, 3)).sum().backward()
    want = x.grad.permute(1, 3, 0, 2, 4).reshape(batch * T, 3 * inner)
    assert torch.isfinite(dqkv.float()).all()
    err = rel_err(dqkv.float().cpu().numpy(), want.cpu().numpy())
    assert err < (1.5e-2 if fmt == "fp16" else 5e-2), err


@pytest.mark.parametrize("cta_group", ["1", "2", "64"])
@pytest.mark.parametrize("M,N,K,splits", [(768, 768, 12608, 16), (3072, 768, 985, 4), (192, 264, 85, 2), (640, 768, 788, 1),
                                          (64, 64, 1000, 7), (128, 2304, 197, 3)])
@pytest.mark.parametrize("fmt", ["fp16", "bf16"])
def test_gemm_tc_wgrad(lib, cta_group, M, N, K, splits, fmt, monkeypatch):
    """dW += X^T dY with X [K, M] and dY [K, N] row-major (MN-major tcgen05 operands, K = the row index,
    ragged K tails zero-filled by TMA), in every tile mode and with split-K."""
    monkeypatch.setenv("VITB200_GEMM_CTA_GROUP", cta_group)
    dt, tdt = DT16[fmt]
    g = torch.Generator().manual_seed(M + N + K)
    X = torch.randn((K, M), generator=g).to(tdt).cuda()
    dY = (torch.randn((K, N), generator=g) / np.sqrt(K)).to(tdt).cuda()
    dW0 = torch.randn((M, N), generator=g).cuda()
    dW = dW0.clone()
    _lib.check(lib.vitb200_gemm_tc_wgrad(stream(), X.data_ptr(), dY.data_ptr(), dW.data_ptr(), M, N, K, splits, dt))
    torch.cuda.synchronize()
    want = X.double().t() @ dY.double() + dW0.double()
    assert (dW.double() - want).abs().max().item() < 3e-4


@pytest.mark.parametrize("rows,dim", [(300, 768), (65, 128), (10, 1280), (33, 192), (1, 64)])
@pytest.mark.parametrize("accumulate", [0, 1])
def test_layernorm_bwd(lib, rows, dim, accumulate):
    dt, tdt = DT16["fp16"]
    g = torch.Generator().manual_seed(rows + dim)
    x = (torch.randn((rows, dim), generator=g) * 2 + 0.5).cuda()
    gamma = (torch.randn(dim, generator=g) * 0.5 + 1).cuda()
    dy = torch.randn((rows, dim), generator=g).to(tdt).cuda()
    dx0 = torch.randn((rows, dim), generator=g).cuda()
    dx = dx0.clone()
    dgamma, dbeta = torch.full((dim,), 2.0, device="cuda"), torch.full((dim,), -1.0, device="cuda")
    _lib.check(lib.vitb200_layernorm_bwd(stream(), dy.data_ptr(), x.data_ptr(), gamma.data_ptr(), dx.data_ptr(),
                                         dgamma.data_ptr(), dbeta.data_ptr(), rows, dim, dt, 1e-6, accumulate))
    torch.cuda.synchronize()
    xr = x.double().requires_grad_(True)
    gr = gamma.double().requires_grad_(True)
    br = torch.zeros(dim, dtype=torch.float64, device="cuda", requires_grad=True)
    (torch.nn.functional.layer_norm(xr, (dim,), gr, br, eps=1e-6) * dy.double()).sum().backward()
    want_dx = xr.grad + (dx0.double() if accumulate else 0)
    assert (dx.double() - want_dx).abs().max().item() < 2e-4 * max(1.0, want_dx.abs().max().item())
    assert (dgamma.double() - 2.0 - gr.grad).abs().max().item() < 1e-4 * max(1.0, gr.grad.abs().max().item())   # accumulates
    assert (dbeta.double() + 1.0 - br.grad).abs().max().item() < 1e-4 * max(1.0, br.grad.abs().max().item())


def _check_grads(eng, variables, cfg, img, dl, pool, tol, logits):
    want_logits, want = vit_torch.vit_vjp(variables, img, dl, pool=pool, **cfg)
    assert np.abs(logits - want_logits).max() < (2e-2 if tol <= 2e-2 else 5e-2)      # forward tolerance of the format (test_gpu_forward.py)
    got = eng.grads()
    ref = flatten_params({"params": want})
    assert set(got) == set(ref)
    errs = {k: rel_err(got[k], ref[k]) for k in ref}
    worst = max(errs, key=errs.get)
    assert errs[worst] < tol, (worst, errs[worst], sorted(errs.items(), key=lambda kv: -kv[1])[:5])
    return errs


@pytest.mark.parametrize("pool", ["cls", "mean"])
@pytest.mark.parametrize("precision", ["fp16", "bf16"])
def test_backward_tiny(pool, precision):
    """Every leaf of a small ViT (ragged everything: T = 17, dim 192, batch 5)."""
    cfg = dict(image_size=32, patch_size=8, num_classes=24, dim=192, depth=2, heads=2, mlp_dim=256)
    variables = perturb_params(init_params(seed=21, **cfg), seed=22)
    img = images_for(cfg, 5, seed=23)
    dl = np.random.default_rng(24).standard_normal((5, 24)).astype(np.float32)
    eng = Engine(precision=precision, max_batch=8, pool=pool, **cfg)
    eng.load_params(variables)
    logits = eng.train_forward(torch.as_tensor(img, device="cuda"))
    eng.backward(torch.as_tensor(dl, device="cuda"))
    torch.cuda.synchronize()
    _check_grads(eng, variables, cfg, img, dl, pool, 2e-2 if precision == "fp16" else 8e-2, logits.cpu().numpy())
    # a second backward of the same forward gives the same gradients (buffers are re-zeroed, nothing accumulates)
    g1 = eng.grads()
    eng.backward(torch.as_tensor(dl, device="cuda"))
    g2 = eng.grads()
    for k in g1:
        assert rel_err(g2[k], g1[k]) < 1e-5, k
    eng.close()


def test_backward_vit_b16_dims():
    """ViT-B/16 geometry (T = 197, dim 768, 12 heads, mlp 3072) at depth 2, batch 4."""
    cfg = dict(C2, depth=2)
    variables = perturb_params(init_params(seed=31, **cfg), seed=32)
    img = images_for(cfg, 4, seed=33)
    dl = np.random.default_rng(34).standard_normal((4, 1000)).astype(np.float32)
    eng = Engine(precision="fp16", max_batch=4, **cfg)
    eng.load_params(variables)
    x = torch.as_tensor(img, device="cuda")
    logits = eng.train_forward(x)
    eng.backward(torch.as_tensor(dl, device="cuda"))
    torch.cuda.synchronize()
    _check_grads(eng, variables, cfg, img, dl, "cls", 2e-2, logits.cpu().numpy())
    assert np.abs(logits.cpu().numpy() - eng.forward(x).cpu().numpy()).max() < 5e-3      # train_forward == forward up to 16-bit rounding
    eng.close()


def test_backward_vit_b16_full_depth_and_full_batch_properties():
    """Full ViT-B/16: (1) all 12 layers against the oracle at batch 2; (2) at BASELINE's batch 256, where the
    float64 oracle would take minutes, the properties a VJP must have: linear in the cotangent, and additive
    over the batch (gradients of 256 images = sum of the gradients of its two halves)."""
    cfg = dict(C2)
    variables = perturb_params(init_params(seed=61, **cfg), seed=62)
    eng = Engine(precision="fp16", max_batch=256, **cfg)
    eng.load_params(variables)
    img = images_for(cfg, 256, seed=63)
    dl = (np.random.default_rng(64).standard_normal((256, 1000)) / 16).astype(np.float32)
    x, d = torch.as_tensor(img, device="cuda"), torch.as_tensor(dl, device="cuda")
    logits = eng.train_forward(x[:2].contiguous())
    eng.backward(d[:2].contiguous())
    _check_grads(eng, variables, cfg, img[:2], dl[:2], "cls", 3e-2, logits.cpu().numpy())
    keys = ["Dense_0/kernel", "pos_embedding", "cls", "Transformer_0/Attention_0/Dense_0/kernel",
            "Transformer_0/FeedForward_5/Dense_1/kernel", "Transformer_0/PreNorm_23/LayerNorm_0/scale",
            "Transformer_0/FeedForward_11/Dense_0/bias", "Dense_1/kernel"]

    def grads_of(lo, hi, scale=1.0):
        eng.train_forward(x[lo:hi].contiguous())
        eng.backward((d[lo:hi] * scale).contiguous())
        return {k: torch.as_tensor(eng.grad_tensor(k)).clone() for k in keys}

    full = grads_of(0, 256)
    a, b = grads_of(0, 128), grads_of(128, 256)
    twice = grads_of(0, 256, 2.0)
    for k in keys:
        ref = full[k].double()
        assert torch.isfinite(ref).all() and ref.abs().max() > 0
        assert ((a[k].double() + b[k].double()) - ref).abs().max() / ref.abs().max() < 5e-3, k     # additive over the batch
        assert (twice[k].double() - 2 * ref).abs().max() / ref.abs().max() < 5e-3, k                # linear in dlogits
    eng.close()


def test_vjp_api_and_loss_scaling():
    """ViT.vjp mirrors jax.vjp(lambda p: v.apply(p, x), params); tiny cotangents survive fp16 through the scaling."""
    cfg = dict(TINY)
    v = ViT(**cfg)
    variables = perturb_params(init_params(seed=41, **cfg), seed=42)
    img = images_for(cfg, 3, seed=43)
    dl = (np.random.default_rng(44).standard_normal((3, cfg["num_classes"])) * 1e-7).astype(np.float32)
    logits, vjp_fn = v.vjp(variables, img)
    assert isinstance(logits, np.ndarray) and logits.shape == (3, cfg["num_classes"])
    grads = vjp_fn(dl)
    _, want = vit_torch.vit_vjp(variables, img, dl, **cfg)
    got, ref = flatten_params(grads), flatten_params({"params": want})
    assert set(got) == set(ref)
    for k in ref:
        assert got[k].shape == ref[k].shape and got[k].dtype == np.float32
        assert rel_err(got[k], ref[k]) < 2e-2, k
    with pytest.raises(ValueError, match="dropout"):                     # rates > 0 need the 'dropout' rng, like apply
        ViT(dropout=0.1, **cfg).vjp(variables, img)
    with pytest.raises(ValueError):
        vjp_fn(dl[:2])


@pytest.mark.parametrize("precision,tol", [("fp16", 2e-2), ("bf16", 8e-2)])
def test_backward_with_dropout_replays_the_forward_masks(precision, tol):
    """vit.py:50,52,83,155 with rates > 0: train_forward drops with the Philox masks of the 'dropout' key and
    the backward replays them (nothing is stored); the oracle differentiates the same dropped forward."""
    cfg = dict(image_size=32, patch_size=8, num_classes=16, dim=128, depth=2, heads=2, mlp_dim=256)
    rate, emb_rate, key = 0.2, 0.1, 0x0123456789ABCDEF
    variables = perturb_params(init_params(seed=71, **cfg), seed=72)
    img = images_for(cfg, 6, seed=73)
    dl = np.random.default_rng(74).standard_normal((6, 16)).astype(np.float32)
    eng = Engine(precision=precision, max_batch=8, dropout=rate, emb_dropout=emb_rate, **cfg)
    eng.load_params(variables)
    eng.set_dropout_key(key)
    x = torch.as_tensor(img, device="cuda")
    inference = eng.forward(x).cpu().numpy()     # BEFORE train_forward: a forward in between voids the kept activations
    logits = eng.train_forward(x).cpu().numpy()
    assert np.abs(logits - inference).max() < 2e-2                             # the same masks as the inference-path dropout
    eng.set_dropout_key(key + 1)                                               # backward must use the FORWARD's key
    eng.backward(torch.as_tensor(dl, device="cuda"))
    want_logits, want = vit_torch.vit_vjp(variables, img, dl, dropout=rate, emb_dropout=emb_rate, dropout_key=key, **cfg)
    assert np.abs(logits - want_logits).max() < (2e-2 if precision == "fp16" else 5e-2)
    got, ref = eng.grads(), flatten_params({"params": want})
    errs = {k: rel_err(got[k], ref[k]) for k in ref}
    worst = max(errs, key=errs.get)
    assert errs[worst] < tol, (worst, errs[worst])
    # and through the module surface
    v = ViT(dropout=rate, emb_dropout=emb_rate, **cfg)
    _, vjp_fn = v.vjp(variables, img, rngs={"dropout": 5}, precision=precision)
    g = flatten_params(vjp_fn(dl))
    assert all(np.isfinite(a).all() for a in g.values()) and set(g) == set(ref)
    eng.close()


def test_backward_more_than_208_tokens():
    """T = 257 (the token count of ViT-H/14 at 224 px): the forward runs the streamed tcgen05 attention and the
    backward the streamed adjoint kernels."""
    cfg = dict(image_size=64, patch_size=4, num_classes=8, dim=128, depth=2, heads=2, mlp_dim=128)
    variables = perturb_params(init_params(seed=81, **cfg), seed=82)
    img = images_for(cfg, 3, seed=83)
    dl = np.random.default_rng(84).standard_normal((3, 8)).astype(np.float32)
    eng = Engine(precision="fp16", max_batch=4, **cfg)
    eng.load_params(variables)
    logits = eng.train_forward(torch.as_tensor(img, device="cuda"))
    eng.backward(torch.as_tensor(dl, device="cuda"))
    _check_grads(eng, variables, cfg, img, dl, "cls", 2e-2, logits.cpu().numpy())
    eng.close()


def test_backward_errors():
    eng = Engine(precision="fp16", max_batch=2, **TINY)
    eng.load_params(init_params(seed=1, **TINY))
    with pytest.raises(VitB200Error, match="train_forward first"):
        eng.backward(torch.zeros((2, TINY["num_classes"]), device="cuda"))
    eng.train_forward(torch.zeros((2, 32, 32, 3), device="cuda"))
    with pytest.raises(VitB200Error, match="batch differs"):
        eng.backward(torch.zeros((1, TINY["num_classes"]), device="cuda"))
    eng.close()
    eng = Engine(precision="fp32", max_batch=1, **TINY)
    eng.load_params(init_params(seed=1, **TINY))
    with pytest.raises(VitB200Error, match="bf16/fp16"):
        eng.train_forward(torch.zeros((1, 32, 32, 3), device="cuda"))
    eng.close()


def test_training_state_is_freed_with_the_model():
    """The activations / gradient workspace of train_forward (GBs at real sizes) must go away with the model."""
    cfg = dict(C2, depth=2)
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    free0 = torch.cuda.mem_get_info()[0]
    eng = Engine(precision="fp16", max_batch=32, **cfg)
    eng.load_params(init_params(seed=1, **cfg))
    x = torch.zeros((32, 224, 224, 3), device="cuda")
    eng.train_forward(x)
    eng.backward(torch.zeros((32, 1000), device="cuda"))
    torch.cuda.synchronize()
    used = free0 - torch.cuda.mem_get_info()[0]
    assert used > 400 << 20                                   # the training state is there
    eng.close()
    del x
    torch.cuda.empty_cache()
    assert free0 - torch.cuda.mem_get_info()[0] < 64 << 20    # and gone (allow allocator / context slack)


@pytest.mark.parametrize("seed", range(8))
def test_backward_config_fuzz(seed):
    """Random small configs with ragged everything (dim not a multiple of 64, odd class counts -> fp32 head path,
    non-square patches, T on both sides of 208, both pools, both formats, batch 1..5) against the autograd oracle."""
    rng = np.random.default_rng(1000 + seed)
    ph, pw = int(rng.choice([4, 8])), int(rng.choice([4, 8, 16]))
    gh, gw = int(rng.integers(1, 9)), int(rng.integers(1, 5))
    if seed in (3, 6):
        ph, pw, gh, gw = 4, 4, 15, 15                                       # T = 226 > 208: streamed attention adjoint
    heads = int(rng.choice([1, 2, 3]))
    dim = int(rng.choice([72, 128, 200]))
    cfg = dict(image_size=(gh * ph, gw * pw), patch_size=(ph, pw), num_classes=int(rng.choice([5, 8, 13, 24])), dim=dim,
               depth=int(rng.integers(1, 3)), heads=heads, mlp_dim=int(rng.choice([64, 136, 256])))
    pool = str(rng.choice(["cls", "mean"]))
    precision = "fp16" if seed % 3 else "bf16"
    batch = int(rng.integers(1, 6))
    variables = perturb_params(init_params(seed=seed, **cfg), seed=seed + 50)
    img = images_for(cfg, batch, seed=seed + 100)
    dl = rng.standard_normal((batch, cfg["num_classes"])).astype(np.float32)
    eng = Engine(precision=precision, max_batch=batch + int(rng.integers(0, 3)), pool=pool, **cfg)
    eng.load_params(variables)
    logits = eng.train_forward(torch.as_tensor(img, device="cuda"))
    eng.backward(torch.as_tensor(dl, device="cuda"))
    _check_grads(eng, variables, cfg, img, dl, pool, 2e-2 if precision == "fp16" else 8e-2, logits.cpu().numpy())
    eng.close()


def test_vjp_fn_refuses_a_stale_forward():
    """vjp_fn closes over the engine of its config (one set of kept activations, a patch workspace shared with
    the inference forward): any forward / reload in between must make it raise, never return wrong gradients."""
    cfg = TINY
    v = ViT(**cfg)
    variables = perturb_params(init_params(seed=1, **cfg), seed=2)
    img = images_for(cfg, 3, seed=3)
    dl = np.ones((3, cfg["num_classes"]), np.float32)
    _, f1 = v.vjp(variables, img)
    g1 = flatten_params(f1(dl))
    _, f2 = v.vjp(variables, img)
    v.apply(variables, images_for(cfg, 3, seed=4))               # overwrites the patch matrix of the kept forward
    with pytest.raises(RuntimeError, match="another forward"):
        f2(dl)
    _, f3 = v.vjp(variables, img)
    _, f4 = v.vjp(variables, images_for(cfg, 3, seed=5))         # a second vjp overwrites the kept activations
    with pytest.raises(RuntimeError, match="another forward"):
        f3(dl)
    g4 = flatten_params(f4(dl))
    assert any(np.abs(g4[k] - g1[k]).max() > 0 for k in g1)
    _, f5 = v.vjp(variables, img)
    g5 = flatten_params(f5(dl))
    for k in g1:     # same inputs again: the same gradients (split-K partial sums meet in any order: not bit-equal)
        np.testing.assert_allclose(g1[k], g5[k], rtol=0, atol=1e-3 * max(1e-6, float(np.abs(g1[k]).max())))
    # the C ABI refuses on its own when an inference forward ran in between
    from vit_flax_b200._lib import VitB200Error
    eng = Engine(precision="fp16", max_batch=3, **cfg)
    eng.load_params(variables)
    x = torch.as_tensor(img, device="cuda")
    eng.train_forward(x)
    eng.forward(x)
    with pytest.raises(VitB200Error, match="train_forward again"):
        eng.backward(torch.as_tensor(dl, device="cuda"))
    eng.close()
    with pytest.raises(VitB200Error, match="at least one transformer layer"):
        e0 = Engine(precision="fp16", max_batch=1, **dict(cfg, depth=0))
        e0.load_params(perturb_params(init_params(seed=1, **dict(cfg, depth=0)), seed=2))
        e0.train_forward(x[:1].contiguous())
