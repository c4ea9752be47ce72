"""Load / save the ``ViT`` params pytree from the on-disk formats a Flax user has.

* ``.msgpack`` -- ``flax.serialization.to_bytes`` / ``msgpack_serialize``: a msgpack map whose leaves
  are ``ExtType(1, packb((shape, dtype_name, raw_bytes)))`` (numpy scalars: code 3), arrays above
  2**30 bytes split into ``{'__msgpack_chunked_array__': True, 'shape': {'0': d0, ...}, 'chunks':
  {'0': ..., ...}}`` (tuples go through flax's ``_tuple_to_dict``).  Restated from the published format
  (flax is not installed here, so no file written by Flax itself could be read back in this image);
  ``save_params`` writes the same layout.
* ``.npz`` -- flat ``'/'``- or ``'.'``-joined paths.
* ``.safetensors`` -- flat paths (``safetensors.flax.save_file`` joins with ``'.'``).

Leaves come back as float32 numpy arrays (bf16 / fp16 checkpoints are widened); the tree is the
nested dict ``ViT.apply`` takes, with or without the top-level ``'params'`` key as stored.
"""
from __future__ import annotations

import os
from typing import Any, Dict, Mapping, Union

import numpy as np

_EXT_NDARRAY, _EXT_NATIVE_COMPLEX, _EXT_NPSCALAR = 1, 2, 3
_MAX_CHUNK = 2 ** 30


def _dtype(name: str) -> np.dtype:
    try:
        return np.dtype(name)
    except TypeError:
        import ml_dtypes  # noqa: F401  (registers bfloat16 & friends with numpy)
        return np.dtype(name)


def _ext_hook(code: int, data: bytes):
    import msgpack
    if code == _EXT_NDARRAY:
        shape, dtype_name, buf = msgpack.unpackb(data, raw=False)
        return np.frombuffer(buf, dtype=_dtype(dtype_name)).reshape(shape)
    if code == _EXT_NPSCALAR:
        shape, dtype_name, buf = msgpack.unpackb(data, raw=False)
        return np.frombuffer(buf, dtype=_dtype(dtype_name)).reshape(shape)[()]
    if code == _EXT_NATIVE_COMPLEX:
        re, im = msgpack.unpackb(data, raw=False)
        return complex(re, im)
    return msgpack.ExtType(code, data)


def _seq(node):
    """A tuple that flax wrote through ``_tuple_to_dict`` (``{'0': a, '1': b}``) or a plain list -> list,
    in index order (keys may come back as str or int)."""
    if isinstance(node, Mapping):
        return [node[k] for k in sorted(node, key=lambda k: int(k))]
    return list(node)


def _unchunk(node):
    if isinstance(node, dict):
        if node.get("__msgpack_chunked_array__"):
            # flax.serialization._chunk: {'__msgpack_chunked_array__': True, 'shape': {'0': d0, ...},
            # 'chunks': {'0': flat chunk, ...}} -- both through _tuple_to_dict; lists are accepted too
            flat = np.concatenate([np.asarray(c).ravel() for c in _seq(node["chunks"])])
            return flat.reshape(tuple(int(d) for d in _seq(node["shape"])))
        return {k: _unchunk(v) for k, v in node.items()}
    return node


def _to_f32(node):
    if isinstance(node, Mapping):
        return {str(k): _to_f32(v) for k, v in node.items()}
    return np.ascontiguousarray(np.asarray(node), dtype=np.float32)


def msgpack_restore(data: bytes) -> Dict[str, Any]:
    """``flax.serialization.msgpack_restore``: bytes -> nested dict of numpy arrays."""
    import msgpack
    return _unchunk(msgpack.unpackb(data, ext_hook=_ext_hook, raw=False, strict_map_key=False))


def _ext_pack(x):
    import msgpack
    if isinstance(x, np.ndarray):
        return msgpack.ExtType(_EXT_NDARRAY, msgpack.packb((x.shape, x.dtype.name, x.tobytes("C")), use_bin_type=True))
    if isinstance(x, np.generic):
        a = np.asarray(x)
        return msgpack.ExtType(_EXT_NPSCALAR, msgpack.packb((a.shape, a.dtype.name, a.tobytes("C")), use_bin_type=True))
    raise TypeError(f"cannot serialise {type(x)}")


def _chunk(node):
    if isinstance(node, Mapping):
        return {str(k): _chunk(v) for k, v in node.items()}
    a = np.asarray(node)
    if a.size * a.dtype.itemsize <= _MAX_CHUNK:
        return a
    per = max(1, _MAX_CHUNK // a.dtype.itemsize)
    flat = a.ravel()
    chunks = {str(i): flat[o:o + per] for i, o in enumerate(range(0, flat.size, per))}
    return {"__msgpack_chunked_array__": True, "shape": {str(i): int(d) for i, d in enumerate(a.shape)}, "chunks": chunks}


def msgpack_serialize(tree: Mapping) -> bytes:
    """``flax.serialization.msgpack_serialize`` of a nested dict of arrays."""
    import msgpack
    return msgpack.packb(_chunk(tree), default=_ext_pack, strict_types=True, use_bin_type=True)


def _unflatten(flat: Mapping[str, Any]) -> Dict[str, Any]:
    tree: Dict[str, Any] = {}
    for key, leaf in flat.items():
        parts = [p for p in key.replace(".", "/").split("/") if p]
        node = tree
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = leaf
    return tree


def load_params(path: Union[str, os.PathLike, bytes]) -> Dict[str, Any]:
    """Read a params pytree; the format is taken from the extension (bytes = msgpack)."""
    if isinstance(path, (bytes, bytearray, memoryview)):
        return _to_f32(msgpack_restore(bytes(path)))
    p = os.fspath(path)
    ext = os.path.splitext(p)[1].lower()
    if ext in (".msgpack", ".flax", ".ckpt"):
        with open(p, "rb") as f:
            return _to_f32(msgpack_restore(f.read()))
    if ext == ".npz":
        with np.load(p) as z:
            return _to_f32(_unflatten({k: z[k] for k in z.files}))
    if ext == ".safetensors":
        from safetensors.numpy import load_file
        return _to_f32(_unflatten(load_file(p)))
    raise ValueError(f"unknown checkpoint format '{ext}' (expected .msgpack, .npz or .safetensors)")


def save_params(variables: Mapping, path: Union[str, os.PathLike]) -> None:
    from .params import flatten_params, leaf_to_numpy
    p = os.fspath(path)
    ext = os.path.splitext(p)[1].lower()
    flat = {k: leaf_to_numpy(v) for k, v in flatten_params(variables).items()}
    has_params = isinstance(variables, Mapping) and "params" in variables
    if ext in (".msgpack", ".flax", ".ckpt"):
        tree = _unflatten(flat)
        with open(p, "wb") as f:
            f.write(msgpack_serialize({"params": tree} if has_params else tree))
    elif ext == ".npz":
        np.savez(p, **{("params/" + k) if has_params else k: a for k, a in flat.items()})
    elif ext == ".safetensors":
        from safetensors.numpy import save_file
        save_file({(("params." if has_params else "") + k.replace("/", ".")): np.ascontiguousarray(a)
                   for k, a in flat.items()}, p)
    else:
        raise ValueError(f"unknown checkpoint format '{ext}'")
