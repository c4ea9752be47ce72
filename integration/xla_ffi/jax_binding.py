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
        jax.ffi.register_ffi_target("vitb200_train_forward", jax.ffi.pycapsule(so.VitB200TrainForward), platform="CUDA")
        jax.ffi.register_ffi_target("vitb200_backward", jax.ffi.pycapsule(so.VitB200Backward), platform="CUDA")
        self._loaded = None

    def apply(self, variables, img, rngs=None):
        jax = self._jax
        if self._loaded is not id(variables):
            self.engine.load_params(variables)      # leaves are read through __array__ (device -> host once)
            self._loaded = id(variables)
        out = jax.ShapeDtypeStruct((img.shape[0], self.cfg["num_classes"]), jax.numpy.float32)
        call = jax.ffi.ffi_call("vitb200_forward", out)
        return call(img.astype(jax.numpy.float32), handle=np.int64(self.engine.handle.value))


    def value_and_vjp(self, variables, img):
        """``jax.vjp(lambda p: self.apply(p, img), variables)``: (logits, vjp_fn).  The activations stay inside the
        library between the two custom calls; the backward call returns every leaf gradient in ONE flat buffer
        (the layout of ``vitb200_grads_buffer``: leaves in ``Engine.param_table()`` order, each padded to a
        multiple of 64 floats), sliced back into the params pytree here."""
        jax, jnp = self._jax, self._jax.numpy
        if self._loaded is not id(variables):
            self.engine.load_params(variables)
            self._loaded = id(variables)
        handle = np.int64(self.engine.handle.value)
        out = jax.ShapeDtypeStruct((img.shape[0], self.cfg["num_classes"]), jnp.float32)
        logits = jax.ffi.ffi_call("vitb200_train_forward", out)(img.astype(jnp.float32), handle=handle)
        table = self.engine.param_table()                       # {flax path: shape}, registry order
        sizes = [int(np.prod(s)) for s in table.values()]
        offsets = np.cumsum([0] + [-(-n // 64) * 64 for n in sizes])

        def vjp_fn(dlogits):
            flat = jax.ffi.ffi_call("vitb200_backward", jax.ShapeDtypeStruct((int(offsets[-1]),), jnp.float32))(
                dlogits.astype(jnp.float32), handle=handle)
            tree = {}
            for (path, shape), off, n in zip(table.items(), offsets, sizes):
                node = tree
                parts = path.split("/")
                for k in parts[:-1]:
                    node = node.setdefault(k, {})
                node[parts[-1]] = flat[off:off + n].reshape(shape)
            return {"params": tree}

        return logits, vjp_fn
