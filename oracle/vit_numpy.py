"""TEST INFRASTRUCTURE ONLY -- numpy restatement of ``vit_flax/vit.py``.

PARITY (see ``oracle/__init__.py``): checked against tests/golden/ref_vit.npz, the outputs of
vit.py itself executed over oracle/flax_shim (jax/flax are absent; the reference has no tests
or golden vectors of its own).  Every function cites the reference lines it restates; library
semantics (flax.linen 0.5 / jax 0.3.13, README.md:837,847) are the published ones:

* ``nn.Dense``      y = x @ kernel (+ bias), kernel ``[in, out]``
* ``nn.LayerNorm``  last axis, eps 1e-6, var = max(0, E[x^2] - E[x]^2)
* ``nn.gelu``       tanh approximation (``approximate=True`` default)
* ``nn.softmax``    exp(x - max) / sum
* ``nn.Dropout``    identity at rate 0 (no rng drawn)

Computation dtype is selectable: float64 (the checker) or float32 (mirrors
what XLA:CPU would do).  Pure numpy; finishes in seconds for the parity sizes.
"""
from __future__ import annotations

import numpy as np

DIM_HEAD = 64  # vit.py:123 (un-annotated class constant, not a ctor field)


def pair(t):
    """vit.py:15-16."""
    return t if isinstance(t, tuple) else (t, t)


def dense(x, p):
    """flax ``nn.Dense`` (call sites vit.py:48,51,68,82,147,165)."""
    y = x @ np.asarray(p["kernel"], dtype=x.dtype)
    if "bias" in p:
        y = y + np.asarray(p["bias"], dtype=x.dtype)
    return y


def layer_norm(x, p, eps=1e-6):
    """flax ``nn.LayerNorm()`` defaults (call sites vit.py:31,163)."""
    mean = x.mean(axis=-1, keepdims=True)
    mean2 = (x * x).mean(axis=-1, keepdims=True)
    var = np.maximum(0.0, mean2 - mean * mean)
    y = (x - mean) / np.sqrt(var + eps)
    return y * np.asarray(p["scale"], dtype=x.dtype) + np.asarray(p["bias"], dtype=x.dtype)


def gelu_tanh(x):
    """``nn.gelu`` default (vit.py:49): tanh approximation."""
    c = float(np.sqrt(2.0 / np.pi))  # python float: weakly typed, keeps x.dtype
    return 0.5 * x * (1.0 + np.tanh(c * (x + 0.044715 * x * x * x)))


def softmax_last(x):
    """``nn.softmax(axis=-1)`` (vit.py:75)."""
    e = np.exp(x - x.max(axis=-1, keepdims=True))
    return e / e.sum(axis=-1, keepdims=True)


def patchify(x, ph, pw):
    """``rearrange(x, 'b (h p1) (w p2) c -> b (h w) (p1 p2 c)')`` (vit.py:146)."""
    b, H, W, c = x.shape
    gh, gw = H // ph, W // pw
    x = x.reshape(b, gh, ph, gw, pw, c).transpose(0, 1, 3, 2, 4, 5)
    return x.reshape(b, gh * gw, ph * pw * c)


def _drop(x, drop, site):
    """``nn.Dropout(rate)(x, deterministic=False)`` with the mask generator of libvitb200
    (oracle/philox.py); ``drop`` = None or (rate, key)."""
    if drop is None or drop[0] == 0.0:
        return x
    from . import philox
    return philox.dropout(x, drop[0], site, drop[1])


def attention(x, p, heads, dim, drop=None, site=0):
    """``Attention.__call__`` (vit.py:62-87)."""
    b, n, _ = x.shape
    inner = DIM_HEAD * heads
    project_out = not (heads == 1 and DIM_HEAD == dim)             # vit.py:65
    scale = DIM_HEAD ** -0.5                                       # vit.py:66
    qkv = dense(x, p["Dense_0"])                                   # vit.py:68 (no bias)
    q, k, v = np.split(qkv, 3, axis=-1)                            # vit.py:69

    def to_heads(t):                                               # vit.py:71
        return t.reshape(b, n, heads, DIM_HEAD).transpose(0, 2, 1, 3)

    q, k, v = to_heads(q), to_heads(k), to_heads(v)
    dots = np.einsum("bhid,bhjd->bhij", q, k) * scale              # vit.py:73
    attn = softmax_last(dots)                                      # vit.py:75
    out = np.einsum("bhij,bhjd->bhid", attn, v)                    # vit.py:78
    out = out.transpose(0, 2, 1, 3).reshape(b, n, inner)           # vit.py:79
    if project_out:
        out = _drop(dense(out, p["Dense_1"]), drop, site)          # vit.py:82-83
    return out


def feed_forward(x, p, drop=None, site=0):
    """``FeedForward.__call__`` (vit.py:47-53): Dense, gelu, Dropout, Dense, Dropout."""
    h = _drop(gelu_tanh(dense(x, p["Dense_0"])), drop, site)        # vit.py:48-50
    return _drop(dense(h, p["Dense_1"]), drop, site + 1)           # vit.py:51-52


def transformer(x, p, depth, heads, dim, drop=None):
    """``Transformer.__call__`` (vit.py:98-112): Residual(PreNorm(.)) pairs.  Dropout sites (the
    numbering of include/vitb200.h): 1+3l after to_out, 2+3l after gelu, 3+3l after the FF output."""
    for l in range(depth):
        ln1 = p[f"PreNorm_{2 * l}"]["LayerNorm_0"]
        x = attention(layer_norm(x, ln1), p[f"Attention_{l}"], heads, dim, drop, 1 + 3 * l) + x   # vit.py:31,39
        ln2 = p[f"PreNorm_{2 * l + 1}"]["LayerNorm_0"]
        x = feed_forward(layer_norm(x, ln2), p[f"FeedForward_{l}"], drop, 2 + 3 * l) + x          # vit.py:31,39
    return x


def vit_forward(variables, images, *, image_size, patch_size, num_classes, dim, depth,
                heads, mlp_dim, pool="cls", dtype=np.float64, return_tokens=False,
                dropout=0.0, emb_dropout=0.0, dropout_key=None):
    """``ViT.__call__`` (vit.py:127-167).  With a rate > 0 the four Dropout sites are applied as Flax
    applies them (deterministic=False), with the masks of oracle/philox.py keyed by ``dropout_key``."""
    p = variables["params"] if "params" in variables else variables
    ih, iw = pair(image_size)
    ph, pw = pair(patch_size)
    assert ih % ph == 0 and iw % pw == 0                           # vit.py:133-134
    assert pool in {"cls", "mean"}                                 # vit.py:137
    x = np.asarray(images, dtype=dtype)
    x = patchify(x, ph, pw)                                        # vit.py:146
    x = dense(x, p["Dense_0"])                                     # vit.py:147
    b, n, _ = x.shape
    cls = np.broadcast_to(np.asarray(p["cls"], dtype=dtype), (b, 1, dim))            # vit.py:151
    x = np.concatenate([cls, x], axis=1)                           # vit.py:152
    x = x + np.asarray(p["pos_embedding"], dtype=dtype)[:, : n + 1]                  # vit.py:153
    if dropout or emb_dropout:
        assert dropout_key is not None, "a dropout rate > 0 needs the 'dropout' rng key"
    x = _drop(x, (emb_dropout, dropout_key), 0)                    # vit.py:155
    x = transformer(x, p["Transformer_0"], depth, heads, dim,
                    (dropout, dropout_key) if dropout else None)   # vit.py:157
    tokens = x
    x = x.mean(axis=1) if pool == "mean" else x[:, 0]              # vit.py:159
    x = layer_norm(x, p["LayerNorm_0"])                            # vit.py:163
    x = dense(x, p["Dense_1"])                                     # vit.py:165
    assert x.shape == (b, num_classes)
    return (x, tokens) if return_tokens else x
