"""TEST INFRASTRUCTURE ONLY -- numpy restatement of ``vit_flax/simple_vit.py`` (parity: tests/golden/ref_simple_vit.npz, the output of simple_vit.py itself over oracle/flax_shim; see
oracle/__init__.py: jax/flax are absent and the reference has no tests).  Reuses the building blocks
of vit_numpy.py; cites the lines of simple_vit.py it follows."""
from __future__ import annotations

import numpy as np

from .vit_numpy import dense, gelu_tanh, pair, softmax_last


def posemb_sincos_2d(h, w, dim, temperature=10000, dtype=np.float64):
    """simple_vit.py:14-25."""
    y, x = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    assert dim % 4 == 0
    omega = np.arange(dim // 4) / (dim // 4 - 1)
    omega = 1.0 / (temperature ** omega)
    y = y.flatten()[:, None] * omega[None, :]
    x = x.flatten()[:, None] * omega[None, :]
    return np.concatenate((np.sin(x), np.cos(x), np.sin(y), np.cos(y)), axis=1).astype(dtype)


def layer_norm_nobias(x, p, eps=1e-5):
    """``nn.LayerNorm(epsilon = 1e-5, use_bias = False)`` (simple_vit.py:41,58,118)."""
    mean = x.mean(axis=-1, keepdims=True)
    var = np.maximum(0.0, (x * x).mean(axis=-1, keepdims=True) - mean * mean)
    return (x - mean) / np.sqrt(var + eps) * np.asarray(p["scale"], dtype=x.dtype)


def attention(x, p, heads, dim_head=64):
    """simple_vit.py:49-73: norm, to_qkv (no bias), softmax attention, to_out (no bias)."""
    b, n, _ = x.shape
    x = layer_norm_nobias(x, p["LayerNorm_0"])
    q, k, v = np.split(dense(x, p["Dense_0"]), 3, axis=-1)
    th = lambda t: t.reshape(b, n, heads, dim_head).transpose(0, 2, 1, 3)
    dots = np.einsum("bhid,bhjd->bhij", th(q), th(k)) * dim_head ** -0.5
    out = np.einsum("bhij,bhjd->bhid", softmax_last(dots), th(v))
    return dense(out.transpose(0, 2, 1, 3).reshape(b, n, heads * dim_head), p["Dense_1"])


def feed_forward(x, p):
    """simple_vit.py:35-45."""
    x = layer_norm_nobias(x, p["LayerNorm_0"])
    return dense(gelu_tanh(dense(x, p["Dense_0"])), p["Dense_1"])


def simple_vit_forward(variables, img, *, image_size, patch_size, num_classes, dim, depth, heads, mlp_dim,
                       channels=3, dim_head=64, dtype=np.float64):
    """``SimpleViT.__call__`` (simple_vit.py:110-134); ``img`` is [B, C, H, W]."""
    p = variables["params"] if "params" in variables else variables
    ih, iw = pair(image_size)
    ph, pw = pair(patch_size)
    assert ih % ph == 0 and iw % pw == 0                                    # simple_vit.py:116
    x = np.asarray(img, dtype=dtype)
    b, c, H, W = x.shape
    gh, gw = H // ph, W // pw
    # 'b c (h p1) (w p2) -> b h w (p1 p2 c)'                                 simple_vit.py:125
    x = x.reshape(b, c, gh, ph, gw, pw).transpose(0, 2, 4, 3, 5, 1).reshape(b, gh, gw, ph * pw * c)
    # Flax names modules at construction (simple_vit.py:117-120 run before :126): LayerNorm_0 / Dense_0 are the
    # head, Dense_1 the patch embedding.  The adoption layout (Dense_0 = patch, Sequential_0/layers_*) that round 1
    # of this repository wrote is accepted too.
    if "Sequential_0" in p:
        patch, head_norm, head_dense = p["Dense_0"], p["Sequential_0"]["layers_0"], p["Sequential_0"]["layers_1"]
    else:
        patch, head_norm, head_dense = p["Dense_1"], p["LayerNorm_0"], p["Dense_0"]
    x = dense(x, patch)                                                     # simple_vit.py:126
    x = x.reshape(b, gh * gw, dim) + posemb_sincos_2d(gh, gw, dim, dtype=dtype)   # simple_vit.py:127-128
    tp = p["Transformer_0"]
    for l in range(depth):                                                  # simple_vit.py:92-95
        x = attention(x, tp[f"Attention_{l}"], heads, dim_head) + x
        x = feed_forward(x, tp[f"FeedForward_{l}"]) + x
    x = x.mean(axis=1)                                                      # simple_vit.py:131
    x = dense(layer_norm_nobias(x, head_norm), head_dense)                  # simple_vit.py:117-120,134
    assert x.shape == (b, num_classes)
    return x
