"""jax.ffi binding of libvitb200 (reference-side stub; needs jax + the adaptor built from
vitb200_xla_ffi.cc -- neither is available in this repository's image).

    from integration.xla_ffi.jax_binding import B200ViT
    v = B200ViT(image_size=224, patch_size=16, num_classes=1000, dim=768, depth=12, heads=12, mlp_dim=3072)
    logits = v.apply(params, img)            # params: the Flax pytree of vit_flax/vit.py; img: jax array on GPU
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class B200ViT:
    """Same constructor fields as vit_flax/vit.py:115-125; ``apply`` runs as one XLA custom call."""

    def __init__(self, *, max_batch=256, precision="fp16", **cfg):
        import jax  # noqa: F401  (fails loudly where jax is absent)
        from vit_flax_b200.engine import Engine
        self._jax = jax
        self.cfg = cfg
        self.engine = Engine(precision=precision, max_batch=max_batch, **cfg)   # owns the C handle
        so = ctypes.CDLL(os.path.join(HERE, "libvitb200_xla.so"))
        jax.ffi.register_ffi_target("vitb200_forward", jax.ffi.pycapsule(so.VitB200Forward), platform="CUDA")
        self._loaded = None

    def apply(self, variables, img, rngs=None):
        jax = self._jax
        if self._loaded is not id(variables):
            self.engine.load_params(variables)      # leaves are read through __array__ (device -> host once)
            self._loaded = id(variables)
        out = jax.ShapeDtypeStruct((img.shape[0], self.cfg["num_classes"]), jax.numpy.float32)
        call = jax.ffi.ffi_call("vitb200_forward", out)
        return call(img.astype(jax.numpy.float32), handle=np.int64(self.engine.handle.value))
