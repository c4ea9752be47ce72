"""TEST INFRASTRUCTURE ONLY -- torch-CPU float32 twin of ``vit_flax/vit.py``.

PARITY: held to the reference run in tests/golden/ref_vit.npz (see ``oracle/__init__.py``).  Written independently of
``vit_numpy.py`` (torch library ops instead of hand-written formulas) so the
two restatements cross-check each other; also the timed CPU baseline
(``bench.py`` ``cpu_baseline`` / ``--impl reference``) because its GEMMs run on
all host cores through MKL/oneDNN, like XLA:CPU's Eigen pool would.

Reference lines restated: vit.py:41-53 (FeedForward), 55-87 (Attention),
89-112 (Transformer), 114-167 (ViT).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

DIM_HEAD = 64  # vit.py:123


def _t(a, dtype):
    return torch.as_tensor(a).to(dtype)


def tree_to_torch(tree, dtype=torch.float32):
    if isinstance(tree, dict) or hasattr(tree, "items"):
        return {k: tree_to_torch(v, dtype) for k, v in tree.items()}
    return _t(tree, dtype)


def _r(x, dt):
    """Round to the 16-bit operand type `dt` and back (no-op when dt is None).  Used to emulate
    WHERE the tensor-core path rounds (GEMM operands and 16-bit activations), so a bf16/fp16 run
    can be checked tightly against "the reference computed with the same operand rounding"."""
    return x if dt is None else x.to(dt).to(x.dtype)


def _ln(x, p):
    # nn.LayerNorm(): eps 1e-6, scale+bias (vit.py:31,163)
    return F.layer_norm(x, (x.shape[-1],), p["scale"], p["bias"], eps=1e-6)


def _drop(x, drop, site):
    """``nn.Dropout(rate)(x, deterministic=False)`` with the Philox masks of libvitb200 (oracle/philox.py);
    ``drop`` = None or (rate, key); x is [..., cols], rows = the flattened leading axes."""
    if drop is None or drop[0] == 0.0:
        return x
    import numpy as np
    from . import philox
    flat_shape = (int(x.numel() // x.shape[-1]), int(x.shape[-1]))
    keep = torch.as_tensor(philox.keep_mask(flat_shape, drop[0], site, drop[1])).view(x.shape)
    inv = float(np.float32(1.0) / (np.float32(1.0) - np.float32(drop[0])))      # the kernels scale in fp32
    return torch.where(keep, x * inv, torch.zeros((), dtype=x.dtype))


def _ln_dense(x, lnp, dense, dt=None, fold=False):
    """``Dense(LayerNorm(x))`` (PreNorm then the first Dense of its fn, vit.py:31 + 48 / 68).  With 16-bit operand
    emulation, ``fold`` selects WHERE the tensor-core path rounds: False = LayerNorm output and kernel (the stand-alone
    LayerNorm kernel); True = the raw x and diag(gamma) W, LayerNorm applied after the product as
    rstd (x16 W'16) - rstd mean c + d with c = 1^T W'16, d = beta^T W + bias (csrc/gemm_tc.cu, LayerNorm fold)."""
    W, bias = dense["kernel"], dense.get("bias")
    if dt is None or not fold:
        y = _r(_ln(x, lnp), dt) @ _r(W, dt)
        return y if bias is None else y + bias
    mean = x.mean(-1, keepdim=True)
    var = (x * x).mean(-1, keepdim=True) - mean * mean                     # flax: E[x^2] - E[x]^2
    rstd = torch.rsqrt(torch.clamp(var, min=0.0) + 1e-6)
    Wp = _r(lnp["scale"][:, None] * W, dt)
    d = lnp["bias"] @ W
    if bias is not None:
        d = d + bias
    return rstd * (_r(x, dt) @ Wp) - rstd * mean * Wp.sum(0) + d


def _attention(x, p, heads, dim, dt=None, drop=None, site=0, pre=None):
    b, n, _ = x.shape
    if pre is not None:      # x is the un-normalised stream, `pre` = (LayerNorm params, fold?)
        qkv = _r(_ln_dense(x, pre[0], p["Dense_0"], dt, pre[1]), dt)
    else:
        qkv = _r(_r(x, dt) @ _r(p["Dense_0"]["kernel"], dt), dt)           # vit.py:68
    q, k, v = qkv.chunk(3, dim=-1)                                         # vit.py:69
    q, k, v = (t.view(b, n, heads, DIM_HEAD).transpose(1, 2) for t in (q, k, v))  # vit.py:71
    if dt is None:
        # vit.py:73-78; SDPA's default scale is 1/sqrt(dim_head) = dim_head ** -0.5
        o = F.scaled_dot_product_attention(q, k, v)
    else:   # un-normalised probabilities are what gets rounded before P @ V on the tensor cores
        s = (q @ k.transpose(-1, -2)) * DIM_HEAD ** -0.5
        e = torch.exp(s - s.amax(dim=-1, keepdim=True))
        o = (_r(e, dt) @ v) / e.sum(dim=-1, keepdim=True)
    o = _r(o.transpose(1, 2).reshape(b, n, heads * DIM_HEAD), dt)          # vit.py:79
    if not (heads == 1 and DIM_HEAD == dim):                               # vit.py:65
        o = _drop(F.linear(o, _r(p["Dense_1"]["kernel"], dt).t(), p["Dense_1"]["bias"]), drop, site)  # vit.py:82-83
    return o


def _ff(x, p, dt=None, drop=None, site=0, pre=None):
    if pre is not None:
        h = F.gelu(_ln_dense(x, pre[0], p["Dense_0"], dt, pre[1]), approximate="tanh")
    else:
        h = F.gelu(_r(x, dt) @ _r(p["Dense_0"]["kernel"], dt) + p["Dense_0"]["bias"], approximate="tanh")  # vit.py:48-49
    h = _drop(h, drop, site)                                                                           # vit.py:50
    return _drop(_r(h, dt) @ _r(p["Dense_1"]["kernel"], dt) + p["Dense_1"]["bias"], drop, site + 1)    # vit.py:51-52


def _vit_forward(params_t, images, *, image_size, patch_size, num_classes, dim, depth,
                heads, mlp_dim, pool="cls", operand_dtype=None, dropout=0.0, emb_dropout=0.0, dropout_key=None,
                ln_fold=False):
    """``params_t``: the ``params`` sub-tree already converted by ``tree_to_torch``.
    ``operand_dtype`` (torch.bfloat16 / torch.float16 / None): emulate 16-bit GEMM operands; ``ln_fold``: round where the
    LayerNorm-folding forward rounds (``_ln_dense``) -- the default inference path of libvitb200 at dropout 0."""
    dt = operand_dtype
    p = params_t["params"] if "params" in params_t else params_t
    ph, pw = (patch_size, patch_size) if not isinstance(patch_size, tuple) else patch_size
    x = torch.as_tensor(images).to(p["cls"].dtype)
    b, H, W, c = x.shape
    # vit.py:146 -- unfold-free patchify through view/permute
    x = x.view(b, H // ph, ph, W // pw, pw, c).permute(0, 1, 3, 2, 4, 5).reshape(b, -1, ph * pw * c)
    x = _r(x, dt) @ _r(p["Dense_0"]["kernel"], dt) + p["Dense_0"]["bias"]  # vit.py:147
    x = torch.cat([p["cls"].expand(b, -1, -1), x], dim=1)                  # vit.py:151-152
    x = x + p["pos_embedding"][:, : x.shape[1]]                            # vit.py:153
    # dropout sites as in include/vitb200.h: 0 emb, 1+3l after to_out, 2+3l after gelu, 3+3l after FF Dense_1
    x = _drop(x, (emb_dropout, dropout_key) if emb_dropout else None, 0)   # vit.py:155
    drop = (dropout, dropout_key) if dropout else None
    tp = p["Transformer_0"]
    for l in range(depth):                                                 # vit.py:108-110
        if dt is not None and ln_fold:
            x = _attention(x, tp[f"Attention_{l}"], heads, dim, dt, drop, 1 + 3 * l, pre=(tp[f"PreNorm_{2 * l}"]["LayerNorm_0"], True)) + x
            x = _ff(x, tp[f"FeedForward_{l}"], dt, drop, 2 + 3 * l, pre=(tp[f"PreNorm_{2 * l + 1}"]["LayerNorm_0"], True)) + x
            continue
        x = _attention(_ln(x, tp[f"PreNorm_{2 * l}"]["LayerNorm_0"]), tp[f"Attention_{l}"], heads, dim, dt, drop, 1 + 3 * l) + x
        x = _ff(_ln(x, tp[f"PreNorm_{2 * l + 1}"]["LayerNorm_0"]), tp[f"FeedForward_{l}"], dt, drop, 2 + 3 * l) + x
    x = x.mean(dim=1) if pool == "mean" else x[:, 0]                       # vit.py:159
    x = _ln(x, p["LayerNorm_0"])                                           # vit.py:163
    return _r(x, dt) @ _r(p["Dense_1"]["kernel"], dt) + p["Dense_1"]["bias"]   # vit.py:165


vit_forward = torch.no_grad()(_vit_forward)


def vit_vjp(variables, images, dlogits, *, dtype=torch.float64, **cfg):
    """The checker of the backward pass: ``jax.vjp(lambda p: ViT.apply(p, img), params)[1](dlogits)``
    restated as torch autograd (float64) through the forward above.  Returns ``(logits, grads)``
    as numpy, ``grads`` a tree shaped like ``variables['params']``."""
    tree = variables["params"] if "params" in variables else variables

    def conv(t):
        if isinstance(t, dict) or hasattr(t, "items"):
            return {k: conv(v) for k, v in t.items()}
        return torch.as_tensor(t).to(dtype).clone().requires_grad_(True)

    p = conv(tree)
    with torch.enable_grad():
        logits = _vit_forward({"params": p}, images, **cfg)
        (logits * torch.as_tensor(dlogits).to(dtype)).sum().backward()

    def grads(t):
        if isinstance(t, dict):
            return {k: grads(v) for k, v in t.items()}
        return (t.grad if t.grad is not None else torch.zeros_like(t)).detach().numpy()

    return logits.detach().numpy(), grads(p)
