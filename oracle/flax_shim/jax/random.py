"""TEST INFRASTRUCTURE ONLY -- `jax.random` stand-in (oracle/flax_shim/README.md).  NOT threefry: a key is a pair of
uint32 words plus the fold-in history, and draws come from numpy's Philox generator seeded with them, so values differ
from JAX's while every call is still a pure function of its key."""
from __future__ import annotations

import zlib

import numpy as _np

from .numpy import wrap


class Key(tuple):
    """(seed words..., fold-ins...) -- hashable, printable, comparable."""
    shape = (2,)


def PRNGKey(seed):
    seed = int(seed)
    return Key((seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))


def fold_in(key, data):
    if isinstance(data, str):
        data = zlib.crc32(data.encode())
    return Key(tuple(key) + (int(data) & 0xFFFFFFFF,))


def split(key, num=2):
    return [fold_in(key, 0x5EED0000 + i) for i in range(num)]


def _gen(key):
    return _np.random.Generator(_np.random.Philox(_np.random.SeedSequence([int(w) for w in key])))


def normal(key, shape=(), dtype=_np.float64):
    return wrap(_gen(key).standard_normal(tuple(shape)).astype(_np.float32).astype(_np.float64))


def uniform(key, shape=(), dtype=_np.float64, minval=0.0, maxval=1.0):
    return wrap(_gen(key).uniform(minval, maxval, tuple(shape)))


def bernoulli(key, p=0.5, shape=()):
    """True with probability p.  tests/golden/make_reference_golden.py replaces this with the mask generator of
    libvitb200 (oracle/philox.py) so that the reference's dropped forward can be compared value for value."""
    return wrap(_gen(key).random(tuple(shape)) < p)


def truncated_normal(key, lower, upper, shape=(), dtype=_np.float64):
    g = _gen(key)
    out = g.standard_normal(tuple(shape))
    bad = (out < lower) | (out > upper)
    while bad.any():
        out[bad] = g.standard_normal(int(bad.sum()))
        bad = (out < lower) | (out > upper)
    return wrap(out)
