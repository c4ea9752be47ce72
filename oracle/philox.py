"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the dropout mask generator of libvitb200
(csrc/ptx.cuh: philox4x32_10 / dropout4), so parity tests can inject the SAME masks into the oracle.

``nn.Dropout(rate)(x, deterministic=False)`` (vit.py:50,52,83,155; flax semantics: keep with
probability 1-rate, kept values scaled by 1/(1-rate)).  JAX's threefry stream cannot be reproduced
without JAX, so bit parity with the reference's masks is impossible by construction; what is
pinned is the SEMANTICS (oracle applies the mask exactly as Flax would) and our generator.
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(quad: np.ndarray, site: int, key: int) -> np.ndarray:
    """quad: uint64 array of counters -> uint32 array [..., 4]."""
    quad = np.asarray(quad, dtype=np.uint64)
    c0 = quad & MASK32
    c1 = quad >> np.uint64(32)
    c2 = np.full_like(c0, np.uint64(site & 0xFFFFFFFF))
    c3 = np.zeros_like(c0)
    k0, k1 = key & 0xFFFFFFFF, (key >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def keep_mask(shape, rate: float, site: int, key: int) -> np.ndarray:
    """Boolean keep-mask of a [rows, cols] activation (flat element index = row * cols + col)."""
    n = int(np.prod(shape))
    words = philox4x32_10(np.arange((n + 3) // 4, dtype=np.uint64), site, key).reshape(-1)[:n]
    t = rate * 4294967296.0
    threshold = 0xFFFFFFFF if t >= 4294967295.0 else int(t)
    return (words >= np.uint32(threshold)).reshape(shape)


def dropout(x: np.ndarray, rate: float, site: int, key: int) -> np.ndarray:
    """Flax ``nn.Dropout`` with our mask: x is [..., cols]; rows are the flattened leading axes."""
    if rate == 0.0:
        return x
    flat = x.reshape(-1, x.shape[-1])
    keep = keep_mask(flat.shape, rate, site, key)
    inv = np.float32(1.0) / (np.float32(1.0) - np.float32(rate))      # the kernels scale in fp32
    return np.where(keep, flat * x.dtype.type(inv), 0).astype(x.dtype).reshape(x.shape)
