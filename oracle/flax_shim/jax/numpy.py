"""TEST INFRASTRUCTURE ONLY -- `jax.numpy` over numpy, float64 (oracle/flax_shim/README.md).

Only what the two reference files call is guaranteed: split, concatenate, mean, einsum (the reference writes its
subscripts with spaces, which numpy ignores like jax does), meshgrid, arange, sin, cos; every other name falls through to
numpy with array results re-wrapped."""
from __future__ import annotations

import numpy as _np


class Array(_np.ndarray):
    """ndarray plus the jax.Array methods the reference uses (simple_vit.py:66: `to_qkv(x).split(3, axis = -1)`)."""

    def split(self, indices_or_sections, axis=0):
        return [wrap(a) for a in _np.split(self, indices_or_sections, axis=axis)]


def wrap(x):
    if isinstance(x, _np.ndarray):
        return x.view(Array)
    if isinstance(x, (list, tuple)):
        return type(x)(wrap(a) for a in x)
    return x


def _lift(fn):
    def call(*args, **kwargs):
        return wrap(fn(*args, **kwargs))
    call.__name__ = getattr(fn, "__name__", "fn")
    return call


def asarray(x, dtype=None):
    a = _np.asarray(x, dtype=dtype)
    if dtype is None and a.dtype in (_np.float32, _np.float16):
        a = a.astype(_np.float64)            # the shim evaluates in float64 (README: what jax_enable_x64 would do)
    return wrap(a)


array = asarray
float32, float64, int32 = _np.float32, _np.float64, _np.int32
ndarray = Array
pi = _np.pi


def __getattr__(name):
    fn = getattr(_np, name)
    return _lift(fn) if callable(fn) and not isinstance(fn, type) else fn
