"""Host-side owner of one ``vitb200_model`` handle (one per device / precision).

PyTorch is plumbing here: it provides the device buffers handed to the C ABI
(``tensor.data_ptr()``) and the stream (``torch.cuda.current_stream()``).  All
compute happens inside ``libvitb200.so``.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np

from . import _lib
from .params import flatten_params, geometry, leaf_to_numpy


def _stream_ptr(torch, device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class Engine:
    """Create -> load_params -> forward.  Not thread-safe (like the C handle)."""

    def __init__(self, *, image_size, patch_size, num_classes, dim, depth, heads, mlp_dim,
                 pool="cls", channels=3, precision="fp16", max_batch=256, device=0,
                 dropout=0.0, emb_dropout=0.0, nchw=False, cls_token=True, ln_eps=0.0):
        import torch  # deferred: plumbing only

        if not torch.cuda.is_available():
            raise RuntimeError("vit_flax_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self._torch = torch
        self.lib = _lib.load()
        ih, iw, ph, pw, n = geometry(image_size, patch_size)
        assert pool in {"cls", "mean"}, "pool type must be either cls (cls token) or mean (mean pooling)"
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS)}, got {precision!r}")
        self.cfg = _lib.Config(
            image_h=ih, image_w=iw, patch_h=ph, patch_w=pw, channels=channels,
            num_classes=num_classes, dim=dim, depth=depth, heads=heads, mlp_dim=mlp_dim,
            pool=_lib.POOL_MEAN if pool == "mean" else _lib.POOL_CLS,
            precision=_lib.PRECISIONS[precision],
            max_batch=max_batch, dropout=float(dropout), emb_dropout=float(emb_dropout),
            flags=(_lib.FLAG_NCHW if nchw else 0) | (0 if cls_token else _lib.FLAG_NO_CLS), ln_eps=float(ln_eps))
        self.precision = precision
        self.max_batch = max_batch
        self.device = torch.device("cuda", device)
        self.tokens = n + (1 if cls_token else 0)
        self.nchw = bool(nchw)
        self.num_classes = num_classes
        self.dim = dim
        self.image_shape = (channels, ih, iw) if nchw else (ih, iw, channels)
        handle = C.c_void_p()
        _lib.check(self.lib.vitb200_create(C.byref(self.cfg), device, C.byref(handle)))
        self.handle = handle
        self._loaded = False
        # Bumped by everything that overwrites what a pending backward reads (any forward, a weight
        # reload, close): ``ViT.vjp`` captures it after ``train_forward`` and ``backward`` refuses a
        # cotangent whose forward is no longer the one held by the handle.
        self.epoch = 0

    # -- params ----------------------------------------------------------------
    def param_table(self) -> Dict[str, tuple]:
        out = {}
        n = _lib.check(self.lib.vitb200_num_params(self.handle))
        for i in range(n):
            path = C.c_char_p()
            shape = (C.c_int64 * 4)()
            nd = _lib.check(self.lib.vitb200_param_info(self.handle, i, C.byref(path), shape))
            out[path.value.decode()] = tuple(shape[:nd])
        return out

    def load_params(self, variables) -> None:
        """Upload a Flax-style params pytree (dict / FrozenDict / Mapping)."""
        self.epoch += 1
        flat = flatten_params(variables)
        table = self.param_table()
        missing = sorted(set(table) - set(flat))
        extra = sorted(set(flat) - set(table))
        if missing or extra:
            raise ValueError(f"params tree does not match the ViT config: missing={missing[:4]} "
                             f"unexpected={extra[:4]}")
        for path, leaf in flat.items():
            a = leaf_to_numpy(leaf)
            shape = (C.c_int64 * max(1, a.ndim))(*a.shape)
            _lib.check(self.lib.vitb200_set_param(self.handle, path.encode(), a.ctypes.data, shape, a.ndim))
        _lib.check(self.lib.vitb200_finalize_params(self.handle, _stream_ptr(self._torch, self.device)))
        self._loaded = True

    def set_dropout_key(self, key: int) -> None:
        """Key of the 'dropout' rng stream; the masks are a pure function of it (like Flax)."""
        _lib.check(self.lib.vitb200_set_dropout_key(self.handle, C.c_uint64(int(key) & 0xFFFFFFFFFFFFFFFF)))

    # -- forward ---------------------------------------------------------------
    def _check_images(self, shape):
        if len(shape) != 4 or tuple(shape[1:]) != self.image_shape:
            raise ValueError(f"expected images [B, {self.image_shape[0]}, {self.image_shape[1]}, "
                             f"{self.image_shape[2]}] ({'NCHW' if self.nchw else 'NHWC, channels-last'}), "
                             f"got {tuple(shape)}")
        if not 1 <= shape[0] <= self.max_batch:
            raise ValueError(f"batch {shape[0]} outside [1, max_batch={self.max_batch}]")

    def forward(self, images, out=None):
        """Device path: ``images`` fp32 CUDA tensor [B,H,W,C]; returns fp32 CUDA logits."""
        torch = self._torch
        self._check_images(images.shape)
        if images.dtype != torch.float32 or not images.is_cuda or not images.is_contiguous():
            raise ValueError("forward expects a contiguous float32 CUDA tensor")
        b = images.shape[0]
        if out is None:
            out = torch.empty((b, self.num_classes), dtype=torch.float32, device=images.device)
        self.epoch += 1
        _lib.check(self.lib.vitb200_forward(self.handle, _stream_ptr(torch, images.device),
                                            images.data_ptr(), b, out.data_ptr()))
        return out

    # -- training --------------------------------------
    def train_forward(self, images, out=None):
        """Forward that keeps every activation the backward pass needs; same arguments as ``forward``.
        (The FeedForward pre-activation is rounded to 16 bits before the GELU here, so the logits can
        differ from ``forward``'s in the last 16-bit digit.)"""
        torch = self._torch
        self._check_images(images.shape)
        if images.dtype != torch.float32 or not images.is_cuda or not images.is_contiguous():
            raise ValueError("train_forward expects a contiguous float32 CUDA tensor")
        b = images.shape[0]
        if out is None:
            out = torch.empty((b, self.num_classes), dtype=torch.float32, device=images.device)
        self.epoch += 1
        _lib.check(self.lib.vitb200_train_forward(self.handle, _stream_ptr(torch, images.device),
                                                  images.data_ptr(), b, out.data_ptr()))
        return out

    def backward(self, dlogits, epoch: Optional[int] = None) -> None:
        """Cotangent of the logits of the last ``train_forward`` (fp32 CUDA tensor [B, classes]) ->
        one fp32 gradient per parameter leaf, read with ``grads()`` / ``grad_tensor(path)``.
        ``epoch``: the value of ``self.epoch`` right after the ``train_forward`` this cotangent belongs
        to; a mismatch means another forward / reload ran on this engine since and is an error."""
        torch = self._torch
        if self.handle is None:
            raise RuntimeError("backward: this engine was closed (a larger batch re-created it); run vjp again")
        if epoch is not None and epoch != self.epoch:
            raise RuntimeError("backward: the engine ran another forward or reloaded its weights since the "
                               "train_forward this cotangent belongs to (activations overwritten); run vjp again")
        if dlogits.dtype != torch.float32 or not dlogits.is_cuda or not dlogits.is_contiguous() or dlogits.dim() != 2 \
                or dlogits.shape[1] != self.num_classes:
            raise ValueError(f"backward expects a contiguous float32 CUDA tensor [B, {self.num_classes}]")
        _lib.check(self.lib.vitb200_backward(self.handle, _stream_ptr(torch, dlogits.device),
                                             dlogits.data_ptr(), int(dlogits.shape[0])))

    def grad_tensor(self, path: str):
        """The gradient of one leaf as a CUDA tensor viewing the library's buffer (overwritten by the
        next ``backward``; clone it to keep it)."""
        torch = self._torch
        shape = self.param_table()[path]
        ptr = C.c_void_p()
        _lib.check(self.lib.vitb200_grad_device(self.handle, path.encode(), C.byref(ptr)))
        n = int(np.prod(shape))

        class _Buf:   # __cuda_array_interface__ view: no copy, no ownership
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (int(ptr.value), False), "version": 2}
        torch.cuda.current_stream(self.device).synchronize()
        return torch.as_tensor(_Buf(), device=self.device).view(*shape)

    def grads_flat(self):
        """Every leaf gradient as one contiguous fp32 CUDA tensor (a view of the library's buffer):
        what a data-parallel step all-reduces (``vit_flax_b200.dist.all_reduce_grads``)."""
        torch = self._torch
        ptr, n = C.c_void_p(), C.c_int64()
        _lib.check(self.lib.vitb200_grads_buffer(self.handle, C.byref(ptr), C.byref(n)))

        class _Buf:
            __cuda_array_interface__ = {"shape": (int(n.value),), "typestr": "<f4", "data": (int(ptr.value), False), "version": 2}
        return torch.as_tensor(_Buf(), device=self.device)

    def grads(self) -> Dict[str, np.ndarray]:
        """{flax path: float32 ndarray} of the last backward pass."""
        out = {}
        for path, shape in self.param_table().items():
            a = np.empty(shape, np.float32)
            _lib.check(self.lib.vitb200_get_grad(self.handle, _stream_ptr(self._torch, self.device),
                                                 path.encode(), a.ctypes.data))
            out[path] = a
        return out

    def forward_host(self, images: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
        """End-to-end path: host fp32 images in, host fp32 logits out (H2D + D2H inside)."""
        images = np.ascontiguousarray(images, dtype=np.float32)
        self._check_images(images.shape)
        b = images.shape[0]
        if out is None:
            out = np.empty((b, self.num_classes), np.float32)
        self.epoch += 1
        _lib.check(self.lib.vitb200_forward_host(self.handle, _stream_ptr(self._torch, self.device),
                                                 images.ctypes.data, b, out.ctypes.data))
        return out

    def submit_host(self, images: np.ndarray, out: np.ndarray) -> None:
        """Pipelined end-to-end path: enqueue H2D + forward + D2H of one batch and return at once
        (at most two jobs in flight; ``wait_host`` completes the oldest).  ``images`` / ``out``
        must be C-contiguous float32 (pinned for real overlap) and stay alive until waited for."""
        if images.dtype != np.float32 or not images.flags.c_contiguous:
            raise ValueError("submit_host expects a C-contiguous float32 array")
        self._check_images(images.shape)
        b = images.shape[0]
        if out.dtype != np.float32 or not out.flags.c_contiguous or out.shape != (b, self.num_classes):
            raise ValueError(f"submit_host expects a C-contiguous float32 out of shape ({b}, {self.num_classes})")
        self.epoch += 1
        _lib.check(self.lib.vitb200_submit_host(self.handle, _stream_ptr(self._torch, self.device),
                                                images.ctypes.data, b, out.ctypes.data))

    def wait_host(self) -> None:
        _lib.check(self.lib.vitb200_wait_host(self.handle))

    def forward_host_stream(self, batches):
        """Generator over host batches -> host logits, in order, two batches in flight: the H2D
        copy of batch k+1 overlaps the forward of batch k.  Logits are written into pinned
        staging arrays that are re-used two batches later: copy them if they must outlive that."""
        torch = self._torch
        outs = [torch.empty((self.max_batch, self.num_classes), dtype=torch.float32).pin_memory().numpy()
                for _ in range(2)]
        pending = []
        k = 0
        for images in batches:
            images = np.ascontiguousarray(images, dtype=np.float32)
            if len(pending) == 2:
                self.wait_host()
                yield pending.pop(0)
            out = outs[k & 1][: images.shape[0]]
            self.submit_host(images, out)
            pending.append(out)
            self._keepalive = images            # the host buffer must outlive the async copy
            k += 1
        while pending:
            self.wait_host()
            yield pending.pop(0)

    def profile_forward(self, images, out=None):
        """One forward with per-launch CUDA events -> {category: (ms, launches)}."""
        torch = self._torch
        self._check_images(images.shape)
        b = images.shape[0]
        if out is None:
            out = torch.empty((b, self.num_classes), dtype=torch.float32, device=images.device)
        n = len(_lib.CATEGORIES)
        ms = (C.c_float * n)()
        cnt = (C.c_int * n)()
        self.epoch += 1
        _lib.check(self.lib.vitb200_profile_forward(self.handle, _stream_ptr(torch, images.device),
                                                    images.data_ptr(), b, out.data_ptr(), ms, cnt))
        return {name: (float(ms[i]), int(cnt[i])) for i, name in enumerate(_lib.CATEGORIES)}

    def tokens_after_transformer(self, batch: int) -> np.ndarray:
        out = np.empty((batch, self.tokens, self.dim), np.float32)
        _lib.check(self.lib.vitb200_debug_tokens(self.handle, _stream_ptr(self._torch, self.device),
                                                 out.ctypes.data, batch))
        return out

    def close(self):
        if getattr(self, "handle", None):
            self.lib.vitb200_destroy(self.handle)
            self.handle = None
            self.epoch += 1

    def __del__(self):  # best effort
        try:
            self.close()
        except Exception:
            pass


def launch_count() -> int:
    return int(_lib.load().vitb200_launch_count())
